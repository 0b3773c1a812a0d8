python -m pytest tests/test_gpu_ddp_qat.py tests/test_gpu_quantsim.py tests/test_gpu_baseline_configs.py -x -q -m gpu 2>&1 | tail -15
python tools/job_timeline.py 8 2>&1 | tail -10
python tools/qat_step.py 2>&1 | tail -6
