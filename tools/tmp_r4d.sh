#!/bin/bash
mkdir -p gpurun_out
L=optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200
timeout 300 python tools/ab_scorer.py new=$L/libb200ret.so p0b1=$L/libb200ret_p0b1.so new2=$L/libb200ret.so > gpurun_out/r4d_ab.jsonl 2> gpurun_out/r4d_ab.err
echo "ab rc=$?"; python -c "
import json
for l in open('gpurun_out/r4d_ab.jsonl'):
    d=json.loads(l); print(d['variant'], d['step_ms'], d['kernel_ms'], d['queries_per_s'], d['same_as_first'])
"; tail -3 gpurun_out/r4d_ab.err
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "approx_prefilter or search_vs_oracle_medium or fused_selection" > gpurun_out/r4d_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r4d_pytest.log
