python -m pytest tests/test_gpu_extract.py tests/test_gpu_pipe.py tests/test_gpu_scale.py -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu --no-full-e2e > gpurun_out/exp_bench.json 2>gpurun_out/exp_bench.err; tail -3 gpurun_out/exp_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/exp_bench.json"))
print("value", d["value"]/1e9, "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"]/1e9)
print(d["roofline"]["stage_ms"], d["roofline"]["frac"])
PY
