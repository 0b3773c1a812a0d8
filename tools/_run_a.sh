set -x
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu_r2c.log 2>&1
tail -5 gpurun_out/pytest_gpu_r2c.log
(time python bench.py > gpurun_out/bench_full_c.json 2> gpurun_out/bench_full_c.err) 2> gpurun_out/bench_full_c.time
tail -3 gpurun_out/bench_full_c.time; tail -5 gpurun_out/bench_full_c.err
python tools/job_timeline.py 8 > gpurun_out/timeline_c.txt 2>&1; cat gpurun_out/timeline_c.txt
