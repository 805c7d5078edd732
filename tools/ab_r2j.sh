B="python bench.py --steps 10 --warmup 3 --secondary 0 --no-cpu-baseline --check 64 --tile-docs 4096"
P=$PWD/optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200
B2R_LIB_PATH=$P/libb200ret_r1.so $B > gpurun_out/r2j_r1.json 2> gpurun_out/r2j_1.err
B2R_LIB_PATH=$P/libb200ret_r1z.so $B > gpurun_out/r2j_r1z.json 2> gpurun_out/r2j_2.err
for f in gpurun_out/r2j_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['parity']['bit_exact_vs_oracle'], d['run'].get('slabs_rank0'))
except Exception as e: print('$f', 'ERR', e)
"; done
tail -3 gpurun_out/r2j_2.err
