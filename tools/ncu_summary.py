#!/usr/bin/env python
"""tools/ncu_summary.py RAW.csv: the handful of metrics of an `ncu --page raw --csv` export that
the profile notes under profiles/ quote, one column per captured launch."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__maximum_warps_per_active_cycle_pct", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = rows[0]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    name_col = hdr.index("Kernel Name")
    print("metric," + ",".join(r[name_col][:60].replace(",", ";") for r in data))
    want = [k for k in KEYS if k in hdr]
    extra = [h for h in hdr if ("pipe_fp64" in h or "issue_stalled" in h and "per_issue_active" in h) and h not in want]
    for k in want + (extra if len(sys.argv) > 2 else []):
        i = hdr.index(k)
        print(k + "," + ",".join(r[i] for r in data))


if __name__ == "__main__":
    main()
