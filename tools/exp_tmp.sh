python -m pytest tests/test_gpu_runlen.py -x -q 2>&1 | tail -8
