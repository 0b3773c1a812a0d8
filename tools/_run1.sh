set -x
python -m pytest tests/test_gpu_stats_multi.py tests/test_gpu_quantsim.py tests/test_gpu_bench_shape.py tests/test_gpu_baseline_configs.py tests/test_quant_analyzer.py tests/test_gpu_distributed.py tests/test_gpu_python_api.py -x -q -m gpu 2>&1 | tail -25
python tools/defer_ab.py 2>&1 | tail -3
python tools/job_phases.py 8 2>&1 | tail -8
