set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r1o.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_r1o.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --bank-schedule 0 > gpurun_out/bench_r1o_nosched.json 2> gpurun_out/bench_r1o_nosched.err; echo rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1o_sched.json 2> gpurun_out/bench_r1o_sched.err; echo rc=$?
cat gpurun_out/bench_r1o_nosched.json gpurun_out/bench_r1o_sched.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity'])"
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 256 --clusters 1 2 4 8 > gpurun_out/cfg_int8_v6.jsonl 2> gpurun_out/cfg_int8_v6.err; echo rc=$?
cat gpurun_out/cfg_int8_v6.jsonl; tail -3 gpurun_out/cfg_int8_v6.err
