#!/usr/bin/env python
"""Hot SASS instructions of one kernel in an .ncu-rep (source page): python tools/ncu_hot.py rep kernel_index [top]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')] + [len(lines)]
blk = lines[starts[kid]:starts[kid + 1]]
print(blk[0][:170])
rows = list(csv.DictReader(io.StringIO("\n".join(blk[1:]))))
rows = [r for r in rows if (r.get("# Samples") or "").isdigit()]
tot = sum(int(r["# Samples"]) for r in rows)
print("total samples", tot, "instructions", len(rows))
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[k]) for r in rows) for k in stalls}
print("stall totals:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"]))[:top]:
    s = int(r["# Samples"])
    why = [f"{k}:{v}" for v, k in sorted(((int(r[k]), k[6:]) for k in stalls), reverse=True)[:3] if v]
    print(f"{i:5d} {s:6d} {100*s/tot:5.1f}%  {r['Source'].strip()[:64]:64s} {' '.join(why)}")
