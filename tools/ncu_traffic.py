#!/usr/bin/env python
"""DRAM traffic per launch of the attention kernels from an `ncu --set full` report of bench.py (B = 256):
    python tools/ncu_traffic.py gpurun_out/prof_bench_attn.ncu-rep profiles/attn_b256_dram_traffic.json
Writes {"mu_attn_bwd": {"dram_bytes_per_launch": ..., ...}, "mu_attn_fwd": {...}} for the 16384-token site (the
launch with the largest grid of each kernel)."""
import csv, io, json, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    v = float(r[col[k]].replace(",", ""))
    u = units[col[k]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "%": 1, "": 1}
    return v * scale.get(u, 1)
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = "mu_attn_bwd" if "attn_bwd_sm100" in name else ("mu_attn_fwd" if "attn_fwd_sm100" in name else None)
    if key is None:
        continue
    grid = num(r, "launch__grid_size")
    rec = {"kernel": name[:80], "grid": grid,
           "dram_bytes_per_launch": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"),
           "dram_read_bytes": num(r, "dram__bytes_read.sum"), "dram_write_bytes": num(r, "dram__bytes_write.sum"),
           "duration_s_under_ncu": num(r, "gpu__time_duration.sum"),
           "tensor_pipe_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
           "xu_pipe_pct": num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed")}
    if key not in out or grid > out[key]["grid"]:
        out[key] = rec
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
