#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, avg, share)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    k = row["Kernel Name"].split("(")[0]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("kernel,launches,total_us,avg_us,share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{t:.1f},{t / n:.1f},{t / tot:.3f}")
