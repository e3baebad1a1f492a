#!/usr/bin/env python
"""Aggregate an ncu gpu__time_duration launch list by kernel: python tools/launch_summary.py launches.csv [top]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    name = row["Kernel Name"]
    k = re.sub(r"\(.*", "", name)
    k = re.sub(r"^void ", "", k)[:110]
    agg[k][0] += 1
    agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1]:9.2f} ms {100 * v[1] / tot:5.1f}% x{v[0]:4d}  {k}")
