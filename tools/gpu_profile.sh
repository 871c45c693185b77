#!/bin/bash
# one gpurun call: the whole -m gpu suite, the launch list of one bench step, --set full of the hot kernels
# usage: tools/gpu_profile.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
python bench.py --steps 5 --warmup 3 --no-full-e2e > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n1.json"))
print("value", d["value"]/1e9, "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"]/1e9)
print(d["roofline"]["stage_ms"], d["roofline"]["frac"])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^(?!.*(at::|unnamed|cub::|thrust::|elementwise)).*" -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-sweep --no-whole > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|encode_kernel|kmerhash_kernel" -s 3 -c 3 -o gpurun_out/${TAG}_extract -f python bench.py --reads 50000 --steps 1 --warmup 0 --no-e2e --no-cpu --no-sweep --no-whole > gpurun_out/${TAG}_ncu_extract.log 2>&1; echo "ncu extract rc=$?"
ncu -i gpurun_out/${TAG}_extract.ncu-rep --page raw --csv > gpurun_out/${TAG}_extract_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
