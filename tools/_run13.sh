mkdir -p gpurun_out
for dg in 0 1 2; do
for cl in 1 2; do
B2R_INT8_DIAG=$dg timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 --clusters $cl --check 0 > gpurun_out/cfg_int8_diag.jsonl 2> gpurun_out/cfg_int8_diag.err; echo rc=$?
python - <<PY
import json
for l in open('gpurun_out/cfg_int8_diag.jsonl'):
    d=json.loads(l); print('diag',$dg,'cluster',d['max_cluster'],'ms',round(d['ms'],3),'tops',round(d['int8_tops'],1))
PY
done
done
