#!/usr/bin/env python3
"""Stall-reason totals and the hottest SASS instructions of one kernel: ncu -i REP --page source --csv | this."""
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
total = sum(int(r[ix["# Samples"]]) for r in data)
print(f"kernel: {rows[0][1]}")
print(f"warp samples: {total}; SASS instructions: {len(data)}; warp instructions executed: {sum(int(r[ix['Instructions Executed']]) for r in data)}")
for h, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
    print(f"  {h:28s} {v:7d} {100.0 * v / max(total, 1):5.1f} %")
print("hottest instructions (samples, SASS, top stall):")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]:
    st = sorted(((h[6:], int(r[ix[h]])) for h in stalls if int(r[ix[h]]) > 0), key=lambda x: -x[1])[:1]
    print(f"  {int(r[ix['# Samples']]):6d}  {r[ix['Source']].strip()[:56]:56s} {st}")
