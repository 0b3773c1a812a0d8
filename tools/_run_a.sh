set -x
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/pytest_gpu_final2.log 2>&1
tail -3 gpurun_out/pytest_gpu_final2.log
python tools/profile_kernels.py pc > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"broadcast_fast_kernel|per_channel_fast_kernel" -s 2 -c 4 -o gpurun_out/r2_run_kernel -f python tools/profile_kernels.py pc > gpurun_out/ncu_run_kernel.log 2>&1
echo rc=$?
ncu -i gpurun_out/r2_run_kernel.ncu-rep --page raw --csv > gpurun_out/r2_run_kernel_raw.csv 2>/dev/null
S="--no-cpu-baseline --no-reference-python --no-other-configs"
python bench.py $S > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err; tail -2 gpurun_out/bench_final2_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_final2_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity']['equals_oracle_checked_golden'])"
