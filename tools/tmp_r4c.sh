#!/bin/bash
mkdir -p gpurun_out
L=optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200
timeout 300 python tools/ab_scorer.py new=$L/libb200ret.so p0b0=$L/libb200ret_p0b0.so p1b0=$L/libb200ret_p1b0.so p0b1=$L/libb200ret_p0b1.so new2=$L/libb200ret.so > gpurun_out/r4c_ab.jsonl 2> gpurun_out/r4c_ab.err
echo "ab rc=$?"; cut -c1-60,150-400 gpurun_out/r4c_ab.jsonl; tail -3 gpurun_out/r4c_ab.err
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "approx_prefilter or search_vs_oracle_medium" > gpurun_out/r4c_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r4c_pytest.log
