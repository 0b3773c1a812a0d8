set -x
python -m pytest tests/test_gpu_quantsim.py tests/test_gpu_bench_shape.py tests/test_quant_analyzer.py tests/test_gpu_percentile.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 --no-kernels --no-cpu-baseline > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err; echo rc=$?; tail -5 gpurun_out/bench_n1_b.err
