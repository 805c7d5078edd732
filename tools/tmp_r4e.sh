#!/bin/bash
mkdir -p gpurun_out
L=optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200
timeout 300 python tools/ab_scorer.py f2=$L/libb200ret.so f0=$L/libb200ret_f0.so f2b=$L/libb200ret.so > gpurun_out/r4e_ab.jsonl 2> gpurun_out/r4e_ab.err
echo "ab rc=$?"; python -c "
import json
for l in open('gpurun_out/r4e_ab.jsonl'):
    d=json.loads(l); print(d['variant'], d['step_ms'], d['kernel_ms'], d['queries_per_s'], d['same_as_first'])
"; tail -3 gpurun_out/r4e_ab.err
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r4e_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r4e_pytest.log
