set -x
(timeout 900 python -m pytest tests/test_reference_unit_tests.py tests/test_reference_kats.py tests/test_reference_python_kats.py tests/test_broadcast_qdq.py tests/test_gpu_range_learning.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_qc_op.py -x -q -m gpu 2>&1 | tail -8) > gpurun_out/pytest_b.log 2>&1
cat gpurun_out/pytest_b.log
python tools/kernel_sweep.py --sizes-mb 64 1024 --out gpurun_out/kernels_f.json > gpurun_out/kernels_f.txt 2>&1
grep -i "per_channel\|blockwise" gpurun_out/kernels_f.txt | grep bf16
