(timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -5)
S="--no-kernels --no-cpu-baseline --no-reference-python --no-other-configs"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 $S > gpurun_out/scale_n2b.json 2> gpurun_out/scale_n2b.err
tail -3 gpurun_out/scale_n2b.err
python - <<'P'
import json
a=json.loads(open('gpurun_out/scale_n1.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/scale_n2b.json').read().strip().splitlines()[-1])
print('N2', b['value'], b['ms_per_step'], b['e2e']['value'], b['strong_scaling'], b['parity'])
print('weak eff vs earlier N1', b['value']/(2*a['value']), 'strong eff', b['strong_scaling']['value']/(2*a['strong_scaling']['value']))
P
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sharded_forward_profile.py 6 2>&1 | grep "job ms"
