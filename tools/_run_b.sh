set -x
(timeout 1200 python -m pytest tests/test_gpu_python_api.py tests/test_gpu_reference_python.py tests/test_install_shim.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_b.log 2>&1
cat gpurun_out/pytest_b.log
python - <<'P'
import bench, json
print(json.dumps(bench.reference_python_leg(), indent=1))
P
AB_DEFER_DROPIN=0 python - <<'P'
import bench, json
r = bench.reference_python_leg(); print("plain class:", r.get("value"), r.get("seconds"), r.get("aimet_b200_launches"))
P
