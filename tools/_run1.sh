set -x
python -m pytest tests/test_gpu_reference_python.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_d.json 2> gpurun_out/bench_n1_d.err; echo rc=$?
tail -3 gpurun_out/bench_n1_d.err
