"""Compact CSV of the metrics that matter from `ncu -i X.ncu-rep --page raw --csv` (stdin) -> stdout."""
import csv
import sys

KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__grid_size launch__block_size
launch__registers_per_thread launch__shared_mem_per_block launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__waves_per_multiprocessor sm__cycles_elapsed.max smsp__cycles_active.avg
sm__warps_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed.sum.pct_of_peak_sustained_elapsed
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__icc_request_hit_rate.pct l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio""".split()
rows = list(csv.reader(sys.stdin))
h, units, data = rows[0], rows[1], rows[2:]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i + 1}" for i in range(len(data))])
w.writerow(["Kernel Name", ""] + [r[h.index("Kernel Name")][:60] for r in data])
for k in KEEP:
    if k in h:
        i = h.index(k)
        w.writerow([k, units[i]] + [r[i] for r in data])
