#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4h_smoke.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r4h_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r4h_pytest_all.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum
B="python bench.py --steps 2 --warmup 3 --secondary 0 --no-cpu-baseline --check 4 --cuda-graph 0"
for w in c2s2 c2s4 c2s8; do
timeout 200 ncu --metrics $M --clock-control none --kernel-name regex:score_approx_kernel --launch-skip 11 --launch-count 1 --csv --log-file gpurun_out/r4h_ncu_approx_$w.csv $B --workload $w > /dev/null 2>&1
echo "ncu $w rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --secondary 0 --no-cpu-baseline --check 64 > gpurun_out/r4h_bench_short.json 2> gpurun_out/r4h_bench_short.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r4h_bench_short.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],4), round(d['value']), json.dumps(d['roofline'])[:1500])
"
