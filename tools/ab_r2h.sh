B="python bench.py --steps 10 --warmup 3 --secondary 0 --no-cpu-baseline --check 64"
P=$PWD/optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab or medium or edge or long_queries or fused_selection or bit_exact or index_file" > gpurun_out/r2h_pytest_slab.log 2>&1; tail -3 gpurun_out/r2h_pytest_slab.log
$B --tile-docs 4096 --slabs 0 > gpurun_out/r2h_t4096_noslab.json 2> gpurun_out/r2h_1.err
$B --tile-docs 4096 > gpurun_out/r2h_t4096_slab.json 2> gpurun_out/r2h_2.err
$B --tile-docs 2048 > gpurun_out/r2h_t2048_slab.json 2> gpurun_out/r2h_3.err
B2R_LIB_PATH=$P/libb200ret_c6.so $B --tile-docs 4096 > gpurun_out/r2h_t4096_slab_c6.json 2> gpurun_out/r2h_4.err
B2R_LIB_PATH=$P/libb200ret_c4.so $B --tile-docs 4096 > gpurun_out/r2h_t4096_slab_c4.json 2> gpurun_out/r2h_5.err
B2R_SLAB_MIN_FRAC=0.125 $B --tile-docs 4096 > gpurun_out/r2h_t4096_slab_f12.json 2> gpurun_out/r2h_6.err
B2R_SLAB_MIN_FRAC=0.5 $B --tile-docs 4096 > gpurun_out/r2h_t4096_slab_f50.json 2> gpurun_out/r2h_7.err
B2R_LIB_PATH=$P/libb200ret_c6.so $B --tile-docs 4096 --slabs 0 > gpurun_out/r2h_t4096_noslab_c6.json 2> gpurun_out/r2h_8.err
for f in gpurun_out/r2h_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['parity']['bit_exact_vs_oracle'], d['run'].get('slabs_rank0'))
except Exception as e: print('$f', 'ERR', e)
"; done
tail -3 gpurun_out/r2h_2.err
