(timeout 1200 python -m pytest tests/test_gpu_python_api.py tests/test_gpu_reference_python.py tests/test_install_shim.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_b.log 2>&1
cat gpurun_out/pytest_b.log
python tools/ref_python_profile.py 2 2>/dev/null | head -1
