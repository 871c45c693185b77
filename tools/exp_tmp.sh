ncu --set full --clock-control none --import-source on -k regex:"scan_kernel" -s 1 -c 1 -o gpurun_out/r02e_scan -f python bench.py --reads 50000 --steps 1 --warmup 0 --no-e2e --no-cpu --no-sweep --no-whole > gpurun_out/r02e_ncu.log 2>&1; echo rc=$?
ncu -i gpurun_out/r02e_scan.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02e_scan_cs.csv 2>/dev/null
ncu -i gpurun_out/r02e_scan.ncu-rep --page raw --csv > gpurun_out/r02e_scan_raw.csv 2>/dev/null
ls -la gpurun_out/r02e*
