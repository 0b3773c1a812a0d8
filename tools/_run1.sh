python -m pytest tests/test_gpu_ddp_qat.py tests/test_gpu_train_graph.py -x -q -m gpu 2>&1 | tail -12
