python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --reads 200000 --config3-reads 600000 > gpurun_out/exp_n2.json 2> gpurun_out/exp_n2.err; echo rc=$?; tail -3 gpurun_out/exp_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/exp_n2.json") if l.startswith("{")][-1])
print("value", d["value"]/1e9, "ms", d["ms_per_step"], d["multi_gpu_parity"]["ok"], d["global_ids_sample_check"]["ok"])
print(json.dumps(d["config3"])[:1500])
PY
