#!/usr/bin/env python
"""Run-to-run determinism of the main kernels: each op twice on identical inputs, max |difference| relative to max |value|.
0 = bit-identical; atomics-ordered reductions (BatchNorm partial sums, split-K weight gradients, dQ reduce-add) may
differ in the last bits."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops  # noqa: E402
from tools.bench_kernels import make  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


def diff(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / a.abs().max().clamp_min(1e-30))


rec = {}
B, N, C = 4, 4096, 64
q, kc, vc, n_keep, keep_idx = make(B, N, C)
o1, l1 = ops.attn_fwd(q, kc, vc, n_keep)
o2, l2 = ops.attn_fwd(q, kc, vc, n_keep)
rec["attn_fwd o / lse"] = [diff(o1, o2), diff(l1, l2)]
d_o = torch.randn_like(o1)
delta = (d_o.float() * o1.float()).sum(-1)
g1 = ops.attn_bwd(q, kc, vc, n_keep, keep_idx, d_o, l1, delta)
g2 = ops.attn_bwd(q, kc, vc, n_keep, keep_idx, d_o, l1, delta)
rec["attn_bwd dq / dk / dv"] = [diff(a, b) for a, b in zip(g1, g2)]
x = torch.randn(8, 64, 64, 64, device=dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
w = torch.randn(128, 64, 3, 3, device=dev, generator=g) * 0.05
y1, s1, _ = ops.conv3x3(x, w, True)
y2, s2, _ = ops.conv3x3(x, w, True)
rec["conv3x3 y / BN partial sums"] = [diff(y1, y2), diff(s1, s2)]
dy = torch.randn_like(y1)
wf, wd = ops.conv_prep_weights(w, True)
rec["conv3x3 dx"] = [diff(ops.conv3x3_bwd_data(dy, wd), ops.conv3x3_bwd_data(dy, wd))]
rec["conv3x3 dw"] = [diff(ops.conv3x3_bwd_weight(x, dy), ops.conv3x3_bwd_weight(x, dy))]
gamma, beta = torch.rand(128, device=dev, generator=g) + 0.5, torch.randn(128, device=dev, generator=g)
b1 = ops.bn_act_fwd(y1, None, gamma, beta, 1e-5, ops.ACT_GELU)
b2 = ops.bn_act_fwd(y1, None, gamma, beta, 1e-5, ops.ACT_GELU)
rec["bn_act_fwd y / mean / rstd"] = [diff(a, b) for a, b in zip(b1[:3], b2[:3])]
print(json.dumps(rec, indent=1))
