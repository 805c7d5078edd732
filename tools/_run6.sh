mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r1t.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/pytest_r1t.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1t.json 2> gpurun_out/bench_r1t.err; echo rc=$?
cat gpurun_out/bench_r1t.json; tail -3 gpurun_out/bench_r1t.err
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1 64 256 1024 > gpurun_out/cfg_int8_v11.jsonl 2> gpurun_out/cfg_int8_v11.err; echo rc=$?
cat gpurun_out/cfg_int8_v11.jsonl; tail -3 gpurun_out/cfg_int8_v11.err
