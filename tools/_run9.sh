mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 --k 1 > gpurun_out/cfg_int8_k1.jsonl 2> gpurun_out/cfg_int8_k1.err; echo rc=$?
cat gpurun_out/cfg_int8_k1.jsonl
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 --k 10 > gpurun_out/cfg_int8_k10.jsonl 2> gpurun_out/cfg_int8_k10.err; echo rc=$?
cat gpurun_out/cfg_int8_k10.jsonl
timeout 600 python tools/bench_configs.py c3 > gpurun_out/cfg_c3_v2.jsonl 2> gpurun_out/cfg_c3_v2.err; echo rc=$?
cat gpurun_out/cfg_c3_v2.jsonl; tail -3 gpurun_out/cfg_c3_v2.err
timeout 600 python tools/bench_configs.py c4 > gpurun_out/cfg_c4_v2.jsonl 2> gpurun_out/cfg_c4_v2.err; echo rc=$?
cat gpurun_out/cfg_c4_v2.jsonl; tail -3 gpurun_out/cfg_c4_v2.err
