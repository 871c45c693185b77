#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_count.py tests/test_gpu_syncasm.py tests/test_cli.py tests/test_gpu_pipe.py tests/test_gpu_golden_pipeline.py tests/test_host_layer.py -m gpu -q > gpurun_out/t_r3.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/t_r3.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print("value", d["value"]/1e9, "ms", d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"]["d2h_bytes_per_step"]/1e9, "full", d["e2e"].get("full_download"))
print("cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"]/1e9)
PY
