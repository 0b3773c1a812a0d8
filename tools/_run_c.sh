python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_final_n8.json 2> gpurun_out/bench_final_n8.err
tail -3 gpurun_out/bench_final_n8.err
python - <<'P'
import json
b=json.loads(open('gpurun_out/bench_final_n8.json').read().strip().splitlines()[-1])
print('N8', b['value'], b['ms_per_step'], b['e2e']['value'], b['strong_scaling'], b['parity'], b['other_configs']['mobilenet_v2_qat'])
P
