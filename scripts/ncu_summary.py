"""Print the key metrics of every kernel in an `ncu --page raw --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, units, data = rows[0], rows[1], rows[2:]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 1
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
ki = h.index('Kernel Name')
idx = {w: h.index(w) for w in want if w in h}
seen = {}
for r in data:
    name = r[ki][:48]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > limit:
        continue
    print('---', name)
    for w, i in idx.items():
        print(f"   {w:82s} {r[i]:>16s} {units[i]}")
