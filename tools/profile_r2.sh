# ncu captures of the scorer (fused-selection epilogue) behind profiles/traffic.json -- run under gpurun, one GPU
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum
B="python bench.py --steps 2 --warmup 3 --secondary 0 --no-cpu-baseline --check 4 --cuda-graph 0"
ncu --set full --clock-control none --import-source on --kernel-name regex:score_tiles_kernel --launch-skip 10 --launch-count 1 -o gpurun_out/r2q_ncu_score_fused_c2 $B > gpurun_out/r2q_ncu_c2.log 2>&1
for w in c2s2 c2s4 c2s8; do
ncu --metrics $M --clock-control none --kernel-name regex:score_tiles_kernel --launch-skip 10 --launch-count 1 --csv --log-file gpurun_out/r2q_ncu_fused_$w.csv $B --workload $w > /dev/null 2>&1
done
ncu --metrics $M --clock-control none --kernel-name regex:score_tiles_kernel --launch-skip 7 --launch-count 1 --csv --log-file gpurun_out/r2q_ncu_fused_c3.csv python tools/bench_configs.py c3 --check 0 > gpurun_out/r2q_c3.log 2>&1
ncu --metrics $M --clock-control none --kernel-name regex:score_tiles_kernel --launch-skip 7 --launch-count 1 --csv --log-file gpurun_out/r2q_ncu_fused_c4.csv python tools/bench_configs.py c4 --check 0 > gpurun_out/r2q_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2q_launches_c2_n1.csv $B > /dev/null 2>&1
ls -la gpurun_out/r2q_*
