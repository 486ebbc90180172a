#!/usr/bin/env python
"""Summarise an ncu raw CSV (ncu -i X.ncu-rep --page raw --csv) into the handful of metrics we track.
usage: ncu_summary.py raw.csv [substring ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_elapsed", "launch__occupancy_limit_registers",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__warp_issue_stalled", "smsp__average_warp", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
extra = sys.argv[2:]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("==", name[:100])
    for i, h in enumerate(hdr):
        if any(k in h for k in KEYS) or any(e in h for e in extra):
            if r[i] not in ("", "0", "n/a"):
                print(f"  {h} [{units[i]}] = {r[i]}")
