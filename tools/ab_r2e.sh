B="python bench.py --steps 10 --warmup 3 --secondary 0 --no-cpu-baseline --check 64"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab or medium or edge or long_queries or fused_selection" > gpurun_out/r2e_pytest_slab.log 2>&1; tail -3 gpurun_out/r2e_pytest_slab.log
$B --slabs 0 > gpurun_out/r2e_ab_noslab.json 2> gpurun_out/r2e_ab_noslab.err
$B > gpurun_out/r2e_ab_slab.json 2> gpurun_out/r2e_ab_slab.err
$B --tile-docs 4096 > gpurun_out/r2e_ab_t4096.json 2> gpurun_out/r2e_ab_t4096.err
B2R_SCORE_RANGES=4 $B > gpurun_out/r2e_ab_slab_r4.json 2> gpurun_out/r2e_ab_slab_r4.err
B2R_SCORE_RANGES=16 $B > gpurun_out/r2e_ab_slab_r16.json 2> gpurun_out/r2e_ab_slab_r16.err
for f in gpurun_out/r2e_ab_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['parity']['bit_exact_vs_oracle'], d['run'].get('slabs_rank0'))
except Exception as e: print('$f', 'ERR', e)
"; done
tail -3 gpurun_out/r2e_ab_slab.err
