mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_tiles -c 3 -o gpurun_out/r1u_bm25 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 --check 0 > gpurun_out/ncu_bm25.log 2>&1; echo rc=$?
ncu -i gpurun_out/r1u_bm25.ncu-rep --page raw --csv > gpurun_out/r1u_bm25_raw.csv 2>/dev/null
ncu -i gpurun_out/r1u_bm25.ncu-rep --page source --csv > gpurun_out/r1u_bm25_source.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1u_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --cuda-graph 0 --check 0 > gpurun_out/ncu_l.log 2>&1; echo rc=$?
ls -la gpurun_out
