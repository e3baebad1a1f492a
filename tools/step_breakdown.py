#!/usr/bin/env python
"""Per-step time by category from an ncu launch list of `bench.py --steps 1 --warmup 1` (first step up to AdamW)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
names = [r["Kernel Name"] for r in rows]
idx = [i for i, n in enumerate(names) if "multi_tensor_apply" in n and "adam" in n.lower()]
# one step = the launches after the previous AdamW group up to and including the last AdamW group
groups = []
for i in idx:
    if groups and i - groups[-1][1] <= 2:
        groups[-1][1] = i
    else:
        groups.append([i, i])
if len(groups) >= 2:
    step = rows[groups[-2][1] + 1: groups[-1][1] + 1]
else:
    step = rows[: idx[-1] + 1] if idx else rows


def classify(n):
    if "attn_bwd" in n or "dq_convert" in n: return "mu attn_bwd"
    if "attn_fwd" in n: return "mu attn_fwd"
    if "mu::qkv" in n or "zero_pad" in n: return "mu qkv projections"
    if "mu::conv_" in n: return "mu conv3x3"
    if "mu::bn_" in n: return "mu bn+act"
    if "mu::residual_ln" in n: return "mu residual LN"
    if "mu::transpose" in n: return "mu transpose"
    if "mu::" in n: return "mu " + re.sub(r"<.*", "", n.split("mu::")[1])[:28]
    if any(k in n for k in ("cutlass", "xmma", "nvjet", "cudnn", "wgrad", "dgrad", "fprop", "gemm", "cublas")): return "lib conv/gemm"
    if "SoftMax" in n or "nll_loss" in n: return "torch CE"
    if "direct_copy" in n or "CatArray" in n: return "torch copy/cat"
    if "multi_tensor" in n: return "torch adamw"
    return "torch " + re.sub(r"<.*", "", n).replace("void at::native::", "").replace("void at::", "")[:44]


cat = collections.defaultdict(lambda: [0.0, 0])
tot = 0.0
for r in step:
    ms = float(r["Metric Value"].replace(",", "")) / 1e6
    c = cat[classify(r["Kernel Name"])]
    c[0] += ms
    c[1] += 1
    tot += ms
print(f"step total {tot:.1f} ms, {len(step)} launches")
for k, v in sorted(cat.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
    print(f"{v[0]:8.2f} ms {100 * v[0] / tot:5.1f}% x{v[1]:4d}  {k}")
