#!/bin/bash
# scratch: NCCL point-to-point channel settings against the exchange / id-return laps
mkdir -p gpurun_out
N=${1:-2}
run() {
    tag=$1; shift
    env "$@" SG_LAPS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-whole --no-sweep --no-e2e --no-config3 \
        > gpurun_out/exp_${tag}.json 2> gpurun_out/exp_${tag}.err
    echo "== $tag rc=$? $(python -c "import json;d=json.loads([l for l in open('gpurun_out/exp_${tag}.json') if l.startswith('{')][-1]);print(round(d['ms_per_step'],2),(d.get('multi_gpu_parity') or {}).get('ok'))")"
    grep "laps dev 0\]" gpurun_out/exp_${tag}.err | grep -E "ids:|exchange:" | tail -2
}
run default X=1
run min32 NCCL_MIN_P2P_NCHANNELS=32
run min32buf NCCL_MIN_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216
run ce NCCL_P2P_USE_CUDA_MEMCPY=1
