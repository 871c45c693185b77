timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02f_bench_n8.json 2> gpurun_out/r02f_bench_n8.err; echo rc=$?; tail -2 gpurun_out/r02f_bench_n8.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02f_bench_n8.json") if l.startswith("{")][-1])
print("value", d["value"]/1e9, "ms", d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["multi_gpu_parity"]["ok"], d["global_ids_sample_check"]["ok"])
print(d["roofline"]["stage_ms"])
for e in d["k_sweep"]: print(e["k"], e["value"]/1e9, e["ms_per_step"])
c=d["config3"]; print("config3", c["value"]/1e9, c["ms_per_step"], c["global_ids_sample_check"]["ok"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02f_bench_n4.json 2> gpurun_out/r02f_bench_n4.err; echo rc=$?
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02f_bench_n4.json") if l.startswith("{")][-1])
print("N4 value", d["value"]/1e9, "ms", d["ms_per_step"], d["multi_gpu_parity"]["ok"], d["global_ids_sample_check"]["ok"])
PY
