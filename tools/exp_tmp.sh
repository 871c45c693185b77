timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_exchange.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-sweep > gpurun_out/exp_n2.json 2> gpurun_out/exp_n2.err; echo rc=$?; tail -2 gpurun_out/exp_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/exp_n2.json") if l.startswith("{")][-1])
print("value", d["value"]/1e9, "ms", d["ms_per_step"], d["multi_gpu_parity"], d["global_ids_sample_check"]["ok"])
sm=d["roofline"]["stage_ms"]; print(sm, sum(sm.values()))
PY
