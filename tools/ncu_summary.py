"""Prints the metrics that matter for this path from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'sm__throughput.avg.pct', 'sm__warps_active.avg.pct_of_peak', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_fp64', 'sm__pipe_fp64_cycles_active', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum ', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'l1tex__t_sector_hit_rate', 'smsp__average_warp', 'issue_stalled',
        'sm__cycles_elapsed.max', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'local_load', 'local_store', 'lmem', 'smsp__inst_executed_op_local',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'sm__warps_active.avg.per_cycle_active', 'smsp__warps_eligible', 'achieved_occupancy', 'sm__inst_executed_pipe_lsu',
        'smsp__inst_executed_op_shared', 'derived__smsp__inst_executed_op_branch', 'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_fma.', 'sm__inst_executed_pipe_xu',
        'smsp__pcsamp_warps_issue_stalled']
seen = set()
for h, u, v in zip(hdr, units, vals):
    if any(w.strip() in h for w in want) and h not in seen:
        seen.add(h)
        print(f"{h} [{u}] = {v}")
