#!/usr/bin/env python3
"""Summarise an .ncu-rep holding many `--set full` captures: one block of key metrics per kernel NAME (the first launch of each, or
the longest with --longest).  Usage (no GPU needed):  python tools/ncu_report_all.py <all_kernels.ncu-rep | its `--page raw --csv` export> [--longest] > profiles/r02_ncu_full_all.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
longest = "--longest" in sys.argv
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
stall = [h for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
ti = hdr.index("gpu__time_duration.sum")
best = {}
for r in rows[2:]:
    name = r[ki].split("(")[0]
    t = float(r[ti].replace(",", "") or 0)
    if name not in best or (longest and t > best[name][0]):
        best[name] = (t, r)
split = sys.argv[sys.argv.index("--split") + 1] if "--split" in sys.argv else None       # --split profiles/r02_ncu_ : one file per kernel too
more = ["lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum", "sm__sass_inst_executed_op_global_ld.sum", "sm__sass_inst_executed_op_global_st.sum"]
import re
for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    if split:
        short = re.sub(r"[^A-Za-z0-9_]+", "_", name.replace("void ", "").replace("unnamed>::", "")).strip("_")
        with open(f"{split}{short}.txt", "w") as fo:
            fo.write(f"{name}  (one `ncu --set full --clock-control none` capture, {'longest' if longest else 'first'} launch of this name; tools/ncu_report_all.py)\n")
            for k in keys + more:
                if k in hdr:
                    i = hdr.index(k); fo.write(f"{k} [{units[i]}] {r[i]}\n")
            for v, h in sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in stall), reverse=True)[:8]:
                fo.write(f"stall {h.split('average_warps_issue_stalled_')[1].split('_per_issue')[0]} {v:.2f} warps/issue\n")
    print(f"== {name}")
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"   {k} [{units[i]}] {r[i]}")
    st = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in stall), reverse=True)[:4]
    for v, h in st:
        print(f"   stall {h.split('average_warps_issue_stalled_')[1].split('_per_issue')[0]} {v:.2f} warps/issue")
