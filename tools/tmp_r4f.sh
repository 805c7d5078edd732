#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "service_save_load or registry or full_size_1m or exchange_merge or batch_pipeline or drop_in_kernels or approx_prefilter" > gpurun_out/r4f_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r4f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4f_bench_n1_full.json 2> gpurun_out/r4f_bench_n1_full.err
echo "bench rc=$?"; tail -3 gpurun_out/r4f_bench_n1_full.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r4f_bench_n1_full.json').read().strip().splitlines()[-1])
    print('step', round(d['ms_per_step'],4), 'value', round(d['value']), 'single', round(d['run']['ms_per_step_one_batch_in_flight'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']), d['parity'])
    for k_,v in (d.get('secondary') or {}).items():
        print(k_, {a:b for a,b in v.items() if a in ('ms_per_step','queries_per_s','parity','search_bm25_queries_per_s','int8_pops','error')})
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('no line', e)
PY
B="python bench.py --steps 2 --warmup 3 --secondary 0 --no-cpu-baseline --check 4 --cuda-graph 0"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:score_approx_kernel --launch-skip 11 --launch-count 1 -o gpurun_out/r4f_ncu_score_approx_c2 $B > gpurun_out/r4f_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r4f_launches_c2_n1.csv $B > /dev/null 2>&1
echo "ncu launches rc=$?"; ls -la gpurun_out/r4f_*
