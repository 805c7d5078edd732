#!/bin/bash
mkdir -p gpurun_out
L=optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bank_schedule or approx_prefilter" > gpurun_out/r4b_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r4b_pytest.log
timeout 300 python tools/ab_scorer.py base=$L/libb200ret.so bank16=$L/libb200ret.so,bank=16 t2=$L/libb200ret.so,B2R_AP_TILES_PER_CTA=2 t8=$L/libb200ret.so,B2R_AP_TILES_PER_CTA=8 u8=$L/libb200ret_u8.so c8=$L/libb200ret_c8.so w8=$L/libb200ret_w8.so w2=$L/libb200ret_w2.so f64=$L/libb200ret.so,prefilter=0 base2=$L/libb200ret.so > gpurun_out/r4b_ab.jsonl 2> gpurun_out/r4b_ab.err
echo "ab rc=$?"; cat gpurun_out/r4b_ab.jsonl; tail -3 gpurun_out/r4b_ab.err
B="python bench.py --steps 2 --warmup 3 --secondary 0 --no-cpu-baseline --check 4 --cuda-graph 0"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:score_approx_kernel --launch-skip 11 --launch-count 1 -o gpurun_out/r4b_ncu_score_approx_c2 $B > gpurun_out/r4b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r4b_ncu.log; ls -la gpurun_out/r4b_*
