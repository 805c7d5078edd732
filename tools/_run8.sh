mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "not int8" > gpurun_out/pytest_r1v.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/pytest_r1v.log
for d in 64 32 16; do
B2R_DENSE_MIN=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1v_d$d.json 2> gpurun_out/bench_r1v_d$d.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_r1v_d$d.json').read()); print($d, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['parity'])"
done
