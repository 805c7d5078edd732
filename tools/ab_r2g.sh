B="python bench.py --steps 10 --warmup 3 --secondary 0 --no-cpu-baseline --check 64"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab or medium or edge or long_queries or fused_selection or bit_exact" > gpurun_out/r2g_pytest_slab.log 2>&1; tail -3 gpurun_out/r2g_pytest_slab.log
$B > gpurun_out/r2g_t2k.json 2> gpurun_out/r2g_t2k.err
$B --slabs 0 > gpurun_out/r2g_t2k_noslab.json 2> gpurun_out/r2g_t2k_noslab.err
B2R_SCORE_RANGES=6 $B > gpurun_out/r2g_t2k_r6.json 2> gpurun_out/r2g_t2k_r6.err
B2R_SCORE_RANGES=10 $B > gpurun_out/r2g_t2k_r10.json 2> gpurun_out/r2g_t2k_r10.err
for f in gpurun_out/r2g_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['parity']['bit_exact_vs_oracle'], d['run'].get('slabs_rank0'))
except Exception as e: print('$f', 'ERR', e)
"; done
tail -3 gpurun_out/r2g_t2k.err
ncu --set full --clock-control none --import-source on --kernel-name regex:score_t2k_kernel --launch-skip 7 --launch-count 1 -o gpurun_out/r2g_ncu_t2k python bench.py --steps 2 --warmup 3 --secondary 0 --no-cpu-baseline --check 8 --cuda-graph 0 > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log | cut -c1-300
