#!/usr/bin/env python
"""Pick the metrics DESIGN.md quotes out of an `ncu -i X.ncu-rep --page raw --csv` export (one kernel launch)."""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__cluster_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main(path, note=""):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    if note:
        print("# " + note)
    print("# kernel: " + vals[col.get("Kernel Name", 4)])
    for k in KEYS:
        if k in col:
            print(f"{k} [{units[col[k]]}] = {vals[col[k]]}")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
