mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "file or save or int8" > gpurun_out/pytest_r1x.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/pytest_r1x.log
for st in 7 6 4; do
B2R_INT8_STAGES=$st timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 --clusters 1 2 4 8 > gpurun_out/cfg_int8_st$st.jsonl 2> gpurun_out/cfg_int8_st$st.err; echo rc=$?
python - <<PY
import json
for l in open('gpurun_out/cfg_int8_st$st.jsonl'):
    d=json.loads(l); print('stages',$st,'cluster',d['max_cluster'],'ms',round(d['ms'],3),'tops',round(d['int8_tops'],1))
PY
done
