#!/usr/bin/env python
"""How well does the attention-forward launch fill the SMs?  (-DMU_FWD_CTALOG=1 build: every CTA logs SM id, entry and
exit clock.)

    python maskunet_b200/build.py --variant maskunet_b200/variant_ctalog.so MU_FWD_CTALOG=1
    MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_ctalog.so python tools/fwd_ctalog.py [--batch 64]
"""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops, _lib  # noqa: E402
from tools.bench_kernels import make  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
B, N, C = a.batch, 16384, 64
q, kc, vc, n_keep, keep_idx = make(B, N, C)
for _ in range(3):
    ops.attn_fwd(q, kc, vc, n_keep)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attn_fwd(q, kc, vc, n_keep); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
lib = _lib.load()
n_cta = B * (N // 128)
buf = (ctypes.c_longlong * (3 * n_cta))()
assert lib.mu_debug_fwd_ctalog(buf, 3 * n_cta) == 0
t = torch.tensor(list(buf), dtype=torch.int64).view(n_cta, 3)
sm, t0, t1 = t[:, 0], t[:, 1], t[:, 2]
dur = (t1 - t0).double()
T = ((n_keep + 127) // 128).cpu().repeat_interleave(N // 128).double()
print(f"launch {ms:.4f} ms; CTAs {n_cta}; CTA duration cycles: median {dur.median():.0f} min {dur.min():.0f} max {dur.max():.0f}")
print(f"cycles per key tile (duration / T): median {(dur / T).median():.0f}, p10 {(dur / T).quantile(0.1):.0f}, p90 {(dur / T).quantile(0.9):.0f}")
busy, span, gaps, ncta = [], [], [], []
for s in sm.unique().tolist():
    m = sm == s
    a0, a1 = t0[m], t1[m]
    lo, hi = a0.min(), a1.max()
    span.append(float(hi - lo)); busy.append(float((a1 - a0).sum())); ncta.append(int(m.sum()))
busy, span = torch.tensor(busy), torch.tensor(span)
print(f"SMs {len(span)}; CTAs per SM {min(ncta)}..{max(ncta)}; per-SM span cycles median {span.median():.0f} max {span.max():.0f}")
print(f"mean resident CTAs per SM over its span: median {(busy / span).median():.3f} min {(busy / span).min():.3f}")
# order of the first CTAs: which linear ids share an SM at the start
first = {}
for lin in range(min(n_cta, 600)):
    first.setdefault(int(sm[lin]), []).append(lin)
print("first CTAs on three SMs:", [first[k][:4] for k in list(first)[:3]])
# start-time spread of the first wave (relative to the earliest start on the same SM)
