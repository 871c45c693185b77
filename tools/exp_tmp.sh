python -m pytest tests/test_gpu_ec.py -x -q -k votes 2>&1 | tail -8
