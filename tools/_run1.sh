set -x
python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -30
