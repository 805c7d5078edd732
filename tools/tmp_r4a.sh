#!/bin/bash
# round-2 session-3 scratch: prefilter parity tests + same-box A/B of the search step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "approx_prefilter or search_vs_oracle_medium" > gpurun_out/r4a_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r4a_pytest.log
for p in 1 0; do
timeout 200 python bench.py --steps 20 --warmup 5 --secondary 0 --no-cpu-baseline --check 64 --prefilter $p > gpurun_out/r4a_pf$p.json 2> gpurun_out/r4a_pf$p.err
echo "bench prefilter=$p rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r4a_pf$p.json').read().strip().splitlines()[-1])
    print('prefilter $p step', round(d['ms_per_step'],4), 'single', round(d['run']['ms_per_step_one_batch_in_flight'],4), 'kernel', round(d['roofline']['kernel_ms'],4), d['roofline']['kernel'], 'e2e', round(d['e2e']['ms_per_step'],4), d['parity'])
except Exception as e:
    print('no line', e)
PY
tail -3 gpurun_out/r4a_pf$p.err
done
