mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:int8_mma -c 2 -o gpurun_out/r1w_int8 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 > gpurun_out/ncu_int8c.log 2>&1; echo rc=$?
ncu -i gpurun_out/r1w_int8.ncu-rep --page raw --csv > gpurun_out/r1w_int8_raw.csv 2>/dev/null
ncu -i gpurun_out/r1w_int8.ncu-rep --page source --csv > gpurun_out/r1w_int8_source.csv 2>/dev/null
