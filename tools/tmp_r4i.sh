#!/bin/bash
mkdir -p gpurun_out
L=optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200
for w in c2s8 c2s4; do
AB_WORKLOAD=$w AB_STEPS=50 timeout 200 python tools/ab_scorer.py t4=$L/libb200ret.so t2=$L/libb200ret.so,B2R_AP_TILES_PER_CTA=2 t1=$L/libb200ret.so,B2R_AP_TILES_PER_CTA=1 t4b=$L/libb200ret.so > gpurun_out/r4i_ab_$w.jsonl 2> gpurun_out/r4i_ab_$w.err
echo "$w rc=$?"; python -c "
import json
for l in open('gpurun_out/r4i_ab_$w.jsonl'):
    d=json.loads(l); print('$w', d['variant'], d['step_ms'], d['kernel_ms'], d['same_as_first'])
"; tail -2 gpurun_out/r4i_ab_$w.err
done
