python -m pytest tests/test_gpu_ec.py tests/test_gpu_syncasm.py tests/test_cli.py tests/test_gpu_survey_kat.py -m gpu -x -q 2>&1 | tail -6
OATK_TIMING=1 python tools/syncasm_run.py --reads 200000 --genome 10000000 --c 30 2> gpurun_out/whole_stages.err | tail -1 | cut -c1-200
grep -n "T::ec" gpurun_out/whole_stages.err | tail -7
