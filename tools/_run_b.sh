(timeout 1200 python -m pytest tests/test_broadcast_qdq.py tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_qc_op.py tests/test_reference_kats.py tests/test_gpu_bounds.py -x -q -m gpu 2>&1 | tail -5)
python tools/kernel_sweep.py --sizes-mb 64 1024 --filter per_channel > gpurun_out/kernels_g.txt 2>&1; grep bf16 gpurun_out/kernels_g.txt
