set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r1q.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_r1q.log
for t in 1 2 4 8; do
B2R_SCORE_TILES_PER_CTA=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1q_t$t.json 2> gpurun_out/bench_r1q_t$t.err; echo rc=$?
done
cat gpurun_out/bench_r1q_t*.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity'])"
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1 64 1024 > gpurun_out/cfg_int8_v8.jsonl 2> gpurun_out/cfg_int8_v8.err; echo rc=$?
cat gpurun_out/cfg_int8_v8.jsonl
timeout 300 ncu --set full --import-source on --clock-control none -k regex:int8_mma -c 2 -o gpurun_out/r1q_int8 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 > gpurun_out/ncu_int8.log 2>&1; echo rc=$?
ncu -i gpurun_out/r1q_int8.ncu-rep --page raw --csv > gpurun_out/r1q_int8_raw.csv 2>/dev/null
ls -la gpurun_out/
