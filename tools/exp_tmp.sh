timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e --no-sweep > gpurun_out/exp_n4.json 2> gpurun_out/exp_n4.err; echo rc=$?; tail -2 gpurun_out/exp_n4.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/exp_n4.json") if l.startswith("{")][-1])
print("N4 value", d["value"]/1e9, "ms", d["ms_per_step"], d["multi_gpu_parity"]["ok"], d["global_ids_sample_check"]["ok"])
print(d["roofline"]["stage_ms"])
PY
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
