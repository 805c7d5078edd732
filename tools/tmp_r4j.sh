#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "approx_prefilter or search_vs_oracle_medium or fused_selection or sharded_merge or full_size_1m" > gpurun_out/r4j_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r4j_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --secondary 0 --no-cpu-baseline --check 1024 > gpurun_out/r4j_bench_short.json 2> gpurun_out/r4j_bench_short.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r4j_bench_short.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],4), round(d['value']), 'single', round(d['run']['ms_per_step_one_batch_in_flight'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['parity'])
"
B="python bench.py --steps 2 --warmup 3 --secondary 0 --no-cpu-baseline --check 4 --cuda-graph 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r4j_launches_c2_n1.csv $B > /dev/null 2>&1
grep "approx_select" gpurun_out/r4j_launches_c2_n1.csv | tail -2 | cut -c1-60,200-400
