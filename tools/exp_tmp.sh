python -m pytest tests/test_gpu_survey_kat.py -x -q 2>&1 | tail -15
