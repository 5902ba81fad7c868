#!/usr/bin/env python3
"""Print the key metrics of an .ncu-rep (one kernel) as 'name unit value' lines: ncu -i REP --page raw --csv | this."""
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr, units, vals = rows[0], rows[1], rows[-1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
keys += [h for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(k, units[i], vals[i])
