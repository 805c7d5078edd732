#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r4g_pytest_multi.log 2>&1
echo "pytest multi rc=$?"; tail -3 gpurun_out/r4g_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r4g_bench_n2_full.json 2> gpurun_out/r4g_bench_n2_full.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/r4g_bench_n2_full.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r4g_bench_n2_full.json').read().strip().splitlines()[-1])
    print('step', round(d['ms_per_step'],4), 'value', round(d['value']), 'single', round(d['run']['ms_per_step_one_batch_in_flight'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']), d['run']['exchange'], d['parity'])
    for k_,v in (d.get('secondary') or {}).items():
        print(k_, {a:b for a,b in v.items() if a in ('ms_per_step','queries_per_s','parity','search_bm25_queries_per_s','int8_pops','error')})
except Exception as e:
    print('no line', e)
PY
