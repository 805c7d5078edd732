B="python bench.py --steps 10 --warmup 3 --secondary 0 --no-cpu-baseline --check 8"
S3=$PWD/optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200/libb200ret_s3.so
$B > gpurun_out/r2f_A_s4.json 2> gpurun_out/r2f_A.err
B2R_SCORE_DIAG=1 $B > gpurun_out/r2f_B_s4_diag.json 2> gpurun_out/r2f_B.err
B2R_LIB_PATH=$S3 $B > gpurun_out/r2f_C_s3.json 2> gpurun_out/r2f_C.err
B2R_LIB_PATH=$S3 B2R_SCORE_DIAG=1 $B > gpurun_out/r2f_D_s3_diag.json 2> gpurun_out/r2f_D.err
B2R_LIB_PATH=$S3 B2R_SCORE_DIAG=1 B2R_SCORE_RANGES=8 $B > gpurun_out/r2f_E_s3_diag_r8.json 2> gpurun_out/r2f_E.err
for f in gpurun_out/r2f_*.json; do python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d['parity']['bit_exact_vs_oracle'], d['run'].get('slabs_rank0'))
except Exception as e: print('$f', 'ERR', e)
"; done
tail -3 gpurun_out/r2f_C.err
