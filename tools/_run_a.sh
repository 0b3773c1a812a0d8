set -x
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/pytest_gpu_final.log 2>&1
tail -5 gpurun_out/pytest_gpu_final.log
(time python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err) 2> gpurun_out/bench_final_n1.time
tail -3 gpurun_out/bench_final_n1.time; tail -3 gpurun_out/bench_final_n1.err
(time python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_final_reference_arm.json 2> gpurun_out/bench_final_reference_arm.err) 2> gpurun_out/bench_final_reference_arm.time
tail -3 gpurun_out/bench_final_reference_arm.time
CMD="python bench.py --steps 4 --warmup 2 --no-kernels --no-cpu-baseline --no-reference-python --no-parity --no-strong --no-other-configs"
$CMD > gpurun_out/ncu_plain_final.log 2>&1 &&
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final2.csv $CMD > gpurun_out/ncu_launches_final2.log 2>&1
echo ncu rc=$?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
