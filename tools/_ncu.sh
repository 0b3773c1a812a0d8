set -x
python -m pytest tests/test_packed.py -q -m gpu 2>&1 | tail -3
CMD="python bench.py --steps 4 --warmup 2 --no-kernels --no-cpu-baseline --no-reference-python --no-parity --no-strong"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on -k regex:hist_multi_kernel -c 2 -o gpurun_out/r2_hist_multi_in_situ -f $CMD > gpurun_out/ncu_full_multi.log 2>&1
echo multi rc=$?
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --nvtx --nvtx-include "timed/" --set full --clock-control none -k regex:"hist_kernel" -s 74 -c 16 -o gpurun_out/r2_hist_single_in_situ -f $CMD > gpurun_out/ncu_full_single.log 2>&1
echo single rc=$?
ls -la gpurun_out | grep ncu-rep
