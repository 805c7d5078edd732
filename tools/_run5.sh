mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "int8" > gpurun_out/pytest_r1s.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_r1s.log
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1 64 256 1024 > gpurun_out/cfg_int8_v10.jsonl 2> gpurun_out/cfg_int8_v10.err; echo rc=$?
cat gpurun_out/cfg_int8_v10.jsonl
