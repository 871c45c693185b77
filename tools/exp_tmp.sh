python -m pytest tests/test_gpu_survey_kat.py -x -q -k reads80k 2>&1 | tail -15
