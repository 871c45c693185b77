#!/bin/bash
# scratch: 8-GPU diagnostics (laps inside sort / exchange / ids, per-rank stage times, NUMA binding on and off)
mkdir -p gpurun_out
{ nproc; lscpu | grep -i -E "numa|socket|model name"; nvidia-smi topo -m; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; } > gpurun_out/r02h_n8_topo.txt 2>&1
N=${1:-8}
SG_LAPS=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-numa-bind --no-cpu --no-whole --no-sweep --no-config3 \
    > gpurun_out/r02h_n${N}_nobind.json 2> gpurun_out/r02h_n${N}_nobind.err
echo "nobind rc=$?"
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-whole \
    > gpurun_out/r02h_bench_n${N}.json 2> gpurun_out/r02h_bench_n${N}.err
echo "bind rc=$?"
grep -h "laps" gpurun_out/r02h_n${N}_nobind.err | sort | uniq -c | sort -rn | head -5 > /dev/null
python - <<PY
import json
for f in ("gpurun_out/r02h_n${N}_nobind.json", "gpurun_out/r02h_bench_n${N}.json"):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line); e = d.get("e2e") or {}
            print(f, "value", round(d["value"] / 1e9, 1), "ms", round(d["ms_per_step"], 2), "e2e", round((e.get("value") or 0) / 1e9, 1), e.get("ms_per_step"), "numa", d.get("numa"), "parity", (d.get("multi_gpu_parity") or {}).get("ok"))
            print("  by_rank", json.dumps(d["roofline"].get("stage_ms_by_rank")))
            print("  c3", json.dumps(d.get("config3"))[:300])
PY
grep "laps dev 0\]" gpurun_out/r02h_n${N}_nobind.err | tail -8
grep "laps dev 5\]" gpurun_out/r02h_n${N}_nobind.err | tail -4
tail -3 gpurun_out/r02h_bench_n${N}.err
