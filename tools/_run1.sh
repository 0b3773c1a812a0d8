python -m pytest tests/test_gpu_qc_op.py tests/test_adaround.py tests/test_packed.py -x -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
