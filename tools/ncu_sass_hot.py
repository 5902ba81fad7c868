#!/usr/bin/env python3
"""Hot spots of one kernel from `ncu --page source --csv --print-source sass`: instructions/stall samples per BAR-delimited section and
the hottest single instructions.  Usage: python tools/ncu_sass_hot.py source_sass.csv [kernel-substring] [min-pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        i = j
        if want not in name: continue
        H = {h: k for k, h in enumerate(hdr)}
        ie, ws, src = H["Instructions Executed"], H["Warp Stall Sampling (All Samples)"], H["Source"]
        num = lambda r, k: int(r[k].replace(",", "") or 0) if r[k].replace(",", "").isdigit() else 0
        ti = sum(num(r, ie) for r in body) or 1; ts = sum(num(r, ws) for r in body) or 1
        print(f"== {name[:90]}  warp-instructions {ti}  stall samples {ts}")
        sec_i = sec_s = 0; sec_start = 0; sec_no = 0
        for k, r in enumerate(body):
            sec_i += num(r, ie); sec_s += num(r, ws)
            if "BAR.SYNC" in r[src] or "EXIT" in r[src] or k == len(body) - 1:
                if sec_i > 0.002 * ti:
                    print(f"  section {sec_no:2d} [{body[sec_start][0]}..{r[0]}] {k - sec_start + 1:4d} instrs: {100 * sec_i / ti:5.1f}% inst {100 * sec_s / ts:5.1f}% stalls  (ends {r[src].split(';')[0].strip()[:40]})")
                sec_no += 1; sec_i = sec_s = 0; sec_start = k + 1
        stall_cols = [h for h in hdr if h.startswith("stall_")]
        for r in body:
            if num(r, ie) > minpct / 100 * ti or num(r, ws) > minpct / 100 * ts:
                top = sorted(((num(r, H[h]), h) for h in stall_cols), reverse=True)[:2]
                print(f"  {r[0]:>6} {100 * num(r, ie) / ti:5.2f}% inst {100 * num(r, ws) / ts:5.2f}% stall  {r[src].split(';')[0].strip()[:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
    else:
        i += 1
