set -x
B="python bench.py --steps 10 --warmup 3 --secondary 0 --no-cpu-baseline --check 64"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab" > gpurun_out/r2c_pytest_slab.log 2>&1; tail -15 gpurun_out/r2c_pytest_slab.log
$B --slabs 0 > gpurun_out/r2c_ab_noslab.json 2> gpurun_out/r2c_ab_noslab.err
$B > gpurun_out/r2c_ab_slab4096.json 2> gpurun_out/r2c_ab_slab4096.err
$B --tile-docs 2048 > gpurun_out/r2c_ab_slab2048.json 2> gpurun_out/r2c_ab_slab2048.err
$B --tile-docs 2048 --slabs 0 > gpurun_out/r2c_ab_noslab2048.json 2> gpurun_out/r2c_ab_noslab2048.err
B2R_SLAB_MIN_FRAC=0.25 $B > gpurun_out/r2c_ab_slab4096_f25.json 2> gpurun_out/r2c_ab_slab4096_f25.err
B2R_SLAB_MIN_FRAC=0.5 $B > gpurun_out/r2c_ab_slab4096_f50.json 2> gpurun_out/r2c_ab_slab4096_f50.err
B2R_SLAB_MIN_FRAC=0.125 $B > gpurun_out/r2c_ab_slab4096_f12.json 2> gpurun_out/r2c_ab_slab4096_f12.err
for f in gpurun_out/r2c_ab_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['parity'], d['run'].get('slabs_rank0'))
"; done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_all.log 2>&1; tail -5 gpurun_out/r2c_pytest_all.log
