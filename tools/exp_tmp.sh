#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_count.py tests/test_gpu_exchange.py tests/test_gpu_scale.py tests/test_gpu_golden_pipeline.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-whole > gpurun_out/exp_n1.json 2> gpurun_out/exp_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/exp_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e9, 2), "ms", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items() if v > 0.01})
for s in d["k_sweep"]: print(s["k"], round(s["value"] / 1e9, 1), {k: round(v, 3) for k, v in s["stage_ms"].items() if v > 0.01})
PY
