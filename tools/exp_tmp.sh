python -m pytest tests/test_gpu_multi.py tests/test_gpu_exchange.py tests/test_gpu_count.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/exp_n2.json 2> gpurun_out/exp_n2.err; echo rc=$?; tail -3 gpurun_out/exp_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/exp_n2.json") if l.startswith("{")][-1])
print("value", d["value"]/1e9, "ms", d["ms_per_step"], d["multi_gpu_parity"]["ok"], d["global_ids_sample_check"]["ok"])
print(d["roofline"]["stage_ms"])
for e in d["k_sweep"]: print(e["k"], e["value"]/1e9, e["ms_per_step"])
PY
