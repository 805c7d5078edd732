set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "int8" > gpurun_out/pytest_r1r.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_r1r.log
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1 64 256 1024 > gpurun_out/cfg_int8_v9.jsonl 2> gpurun_out/cfg_int8_v9.err; echo rc=$?
cat gpurun_out/cfg_int8_v9.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/r1r_int8_launches.csv python tools/bench_configs.py int8 --docs 2000000 --queries 1024 > gpurun_out/ncu_int8b.log 2>&1; echo rc=$?
