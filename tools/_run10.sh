mkdir -p gpurun_out
for t in 4 1 2; do
B2R_SCORE_TILES_PER_CTA=$t timeout 600 python tools/bench_configs.py c3 > gpurun_out/cfg_c3_t$t.jsonl 2> gpurun_out/cfg_c3_t$t.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/cfg_c3_t$t.jsonl').read()); print('tiles/cta', $t, d['ms_per_batch'], d['ms_per_batch_plain_path'], d['queries_per_s'])"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
