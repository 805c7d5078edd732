set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r1p.log 2>&1; echo pytest rc=$?
tail -5 gpurun_out/pytest_r1p.log
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg
for s in 0 1; do
timeout 300 ncu --metrics $M -k regex:score_tiles --clock-control none -c 3 --csv --log-file gpurun_out/r1p_ncu_sched$s.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cuda-graph 0 --check 0 --bank-schedule $s > gpurun_out/ncu_sched$s.log 2>&1; echo rc=$?
done
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1 16 64 256 1024 --clusters 1 --fused-modes 0 2 > gpurun_out/cfg_int8_v7.jsonl 2> gpurun_out/cfg_int8_v7.err; echo rc=$?
cat gpurun_out/cfg_int8_v7.jsonl; tail -3 gpurun_out/cfg_int8_v7.err
timeout 300 python tools/bench_configs.py int8 --docs 2000000 --queries 1024 --clusters 2 4 8 --fused-modes 2 > gpurun_out/cfg_int8_v7b.jsonl 2> gpurun_out/cfg_int8_v7b.err; echo rc=$?
cat gpurun_out/cfg_int8_v7b.jsonl
