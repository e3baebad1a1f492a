#!/usr/bin/env python
"""HBM-bound kernels of one training step from an `ncu --set full` report: per kernel variant the launch with the most DRAM
bytes -- duration, DRAM bytes read + written, achieved GB/s and its fraction of the measured copy bandwidth.
    python tools/ncu_hbm_table.py report.ncu-rep [out.txt]"""
import csv, io, json, os, re, subprocess, sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "%": 1, "": 1}


def num(r, k):
    v = r[col[k]].replace(",", "")
    return float(v) * SCALE.get(units[col[k]].lower(), 1) if v else 0.0


peak = 6537.0
try:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = float(json.load(open(os.path.join(here, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
best = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("mu::", "")
    t = num(r, "gpu__time_duration.sum")
    byt = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    rec = (byt, t, num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum"),
           num(r, "launch__registers_per_thread"), num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
           r[col["Grid Size"]] if "Grid Size" in col else "")
    if name not in best or byt > best[name][0]:
        best[name] = rec
lines = ["# HBM-bound kernels inside one benchmark step (ncu --set full --clock-control none, B = 256): the launch of each kernel",
         "# variant that moves the most DRAM bytes; peak = measured copy bandwidth %.0f GB/s (MEASURED_PEAKS.json)" % peak,
         "%-64s %9s %9s %9s %9s %6s %5s %6s" % ("kernel", "us", "read MB", "write MB", "GB/s", "frac", "regs", "warps%")]
for name, (byt, t, rd, wr, regs, warps, grid) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    if t <= 0:
        continue
    gbs = byt / t / 1e9
    lines.append("%-64s %9.1f %9.1f %9.1f %9.0f %6.2f %5.0f %6.1f" % (name[:64], t * 1e6, rd / 1e6, wr / 1e6, gbs, gbs / peak, regs, warps))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
