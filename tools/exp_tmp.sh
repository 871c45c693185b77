python -m pytest tests/test_gpu_count.py tests/test_gpu_scale.py tests/test_gpu_exchange.py -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-whole > gpurun_out/exp_bench.json 2>gpurun_out/exp_bench.err; tail -3 gpurun_out/exp_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/exp_bench.json"))
print("value", d["value"]/1e9, "ms", d["ms_per_step"])
print(d["roofline"]["stage_ms"], d["roofline"]["stage_launches"])
for e in d["k_sweep"]: print(e["k"], e["value"]/1e9, e["stage_ms"])
PY
