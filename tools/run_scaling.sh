#!/bin/bash
# usage: tools/run_scaling.sh N TAG   -- bench.py (C2, strong scaling) and the C3 sharded measurement on N GPUs of one node
N=$1; TAG=$2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo rc=$?
  timeout 600 python tools/bench_configs.py c3 > gpurun_out/cfg_c3_${TAG}_n1.jsonl 2> gpurun_out/cfg_c3_${TAG}_n1.err; echo rc=$?
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo rc=$?
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_configs.py c3 > gpurun_out/cfg_c3_${TAG}_n$N.jsonl 2> gpurun_out/cfg_c3_${TAG}_n$N.err; echo rc=$?
fi
tail -c 2500 gpurun_out/bench_${TAG}_n$N.json; echo; tail -2 gpurun_out/bench_${TAG}_n$N.err
cat gpurun_out/cfg_c3_${TAG}_n$N.jsonl; tail -2 gpurun_out/cfg_c3_${TAG}_n$N.err
