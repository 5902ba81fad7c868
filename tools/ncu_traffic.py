#!/usr/bin/env python3
"""DRAM traffic per bench step of every frame-construction stage from the `--page raw --csv` export of the `ncu --set full` captures
(tools/profile_r02.sh): dram__bytes_read.sum + dram__bytes_write.sum, averaged per distinct (kernel, grid, instruction count) launch shape (one launch
of each shape per step), summed per stage.  Writes the JSON bench.py reads (profiles/r02_ncu_traffic.json) and prints a table.
Usage: python tools/ncu_traffic.py gpurun_out/r02/<tag>/fc_kernels_raw.csv [chain_kernels_raw.csv] > profiles/r02_ncu_traffic.json"""
import csv, io, json, sys
STAGE = {"level_tile_kernel": "pyramid", "resize_level": "pyramid", "blur_level": "blur", "fast_strips": "fast", "fast_cells": "fast", "cand_scan": "compact",
         "cand_gather": "compact", "quadtree_kernel": "quadtree", "sel_pack": "quadtree", "describe": "describe", "depth_project": "depth_project",
         "depth_resolve_dilate": "depth_resolve_dilate", "depth_gather": "depth_gather"}
def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
shapes = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    H = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[H["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        num = lambda k: float(r[H[k]].replace(",", "") or 0) * unit_scale(units[H[k]])
        inst = float(r[H["smsp__inst_executed.sum"]].replace(",", "") or 0)
        key = (name, r[H["launch__grid_size"]] + f", {float(f'{inst:.2g}'):.0f} warp-instr")      # persistent kernels share a grid size across levels
        s = shapes.setdefault(key, {"n": 0, "rd": 0.0, "wr": 0.0, "us": 0.0})
        s["n"] += 1; s["rd"] += num("dram__bytes_read.sum"); s["wr"] += num("dram__bytes_write.sum")
        s["us"] += float(r[H["gpu__time_duration.sum"]].replace(",", "") or 0) * {"us": 1, "ns": 1e-3, "ms": 1e3}.get(units[H["gpu__time_duration.sum"]], 1)
stages, detail = {}, {}
for (name, grid), s in sorted(shapes.items()):
    mb = (s["rd"] + s["wr"]) / s["n"] / 1e6
    detail[f"{name}[grid {grid}]"] = {"captures": s["n"], "dram_read_MB": round(s["rd"] / s["n"] / 1e6, 3), "dram_write_MB": round(s["wr"] / s["n"] / 1e6, 3),
                                      "us_under_ncu": round(s["us"] / s["n"], 2)}
    for pat, st in STAGE.items():
        if pat in name:
            stages[st] = stages.get(st, 0.0) + mb
            break
out = {k: round(v, 3) for k, v in stages.items()}
out["_per_launch_shape"] = detail
out["_note"] = "MB of DRAM traffic (read + write) per 32-frame step and stage; ncu --set full, cold caches, serialised launches"
print(json.dumps(out, indent=1))
