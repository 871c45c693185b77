#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t_multi.log 2>&1; echo "multi rc=$?"; tail -30 gpurun_out/t_multi.log
python -m pytest tests/test_cli.py -m gpu -q > gpurun_out/t_cli.log 2>&1; echo "cli rc=$?"; tail -30 gpurun_out/t_cli.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --reads 250000 > gpurun_out/bench_n2_small.json 2> gpurun_out/bench_n2_small.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n2_small.json; tail -5 gpurun_out/bench_n2_small.err
