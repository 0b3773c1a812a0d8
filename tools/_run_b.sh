(timeout 600 python -m pytest tests/test_gpu_range_learning.py -x -q -m gpu 2>&1 | tail -3)
python tools/kernel_sweep.py --sizes-mb 64 1024 --filter range_learning > gpurun_out/kernels_h.txt 2>&1; cat gpurun_out/kernels_h.txt
