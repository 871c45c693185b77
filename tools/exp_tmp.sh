python bench.py --steps 3 --warmup 3 --no-full-e2e > gpurun_out/exp_bench.json 2>gpurun_out/exp_bench.err; echo rc=$?; tail -5 gpurun_out/exp_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/exp_bench.json"))
print("value", d["value"]/1e9, "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"]/1e9)
print(json.dumps(d["k_sweep"])[:1500])
print(json.dumps(d["whole_command"])[:3000])
PY
