python -m pytest tests/test_gpu_extract.py tests/test_gpu_scale.py tests/test_gpu_survey_kat.py tests/test_gpu_pipe.py -x -q -k "not reads80k" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-whole 2>gpurun_out/exp_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value', d['value']/1e9, 'ms', d['ms_per_step']); print(d['roofline']['stage_ms'])
for e in d['k_sweep']: print(e['k'], e['value']/1e9, e['stage_ms']['scan'])"
