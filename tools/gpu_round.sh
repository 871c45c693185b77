#!/bin/bash
# one gpurun call: extract parity tests, smoke launch list, bench on both input classes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_extract.py -x -q > gpurun_out/t_extract.log 2>&1; echo "extract tests rc=$?"; tail -5 gpurun_out/t_extract.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
grep -E "scan_kernel|scan_exact|encode_kernel" gpurun_out/smoke_launches.csv | cut -d, -f5,12- | head
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err; tail -c 1500 gpurun_out/bench_uniform.json
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload repeats > gpurun_out/bench_repeats.json 2> gpurun_out/bench_repeats.err; tail -c 1500 gpurun_out/bench_repeats.json
