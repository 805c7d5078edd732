"""Print the judged counters of an `ncu --page raw --csv` export (one block per kernel)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'memory_l1_wavefronts_shared_ideal',
        'derived__memory_l1_wavefronts_shared_excessive', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'smsp__inst_executed.sum', 'smsp__inst_executed_pipe_fp64.sum',
        'smsp__cycles_active.avg', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:110])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:75s} {r[i]:>20s} {units[i]}")
