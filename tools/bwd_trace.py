#!/usr/bin/env python
"""Timeline of ONE attention-backward CTA (trace build of the library, -DMU_BWD_TRACE=1):

    python maskunet_b200/build.py --variant maskunet_b200/build_variant_trace.so MU_BWD_TRACE=1
    MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_trace.so python tools/bwd_trace.py

Prints, per query tile, the SM-clock offsets of the pipeline events of CTA (0, 0, 0) relative to the tile's
'softmax: s_full' event, and the tile-to-tile period."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops, _lib  # noqa: E402
from tools.bench_kernels import make  # noqa: E402

NAMES = ["tma:issue", "mma:qdo_full", "mma:sdp_free", "mma:pds_full", "mma:dVdK issued", "mma:dq_free", "mma:acc issued",
         "sm:wait s", "sm:s_full", "sm:in regs", "sm:bufs free", "sm:published", "dq:dq_full", "dq:stage free",
         "dq:reduce issued", "sm:chunk0", "sm:chunk1", "sm:wait st", "sm:fence", "sm2:s_full", "sm2:bufs free",
         "sm2:published", "-", "-"]
B, N, C = 16, 16384, 64
q, kc, vc, n_keep, keep_idx = make(B, N, C)
o, lse = ops.attn_fwd(q, kc, vc, n_keep)
d_o = torch.randn_like(o)
delta = (d_o.float() * o.float()).sum(-1)
for _ in range(2):
    ops.attn_bwd(q, kc, vc, n_keep, keep_idx, d_o, lse, delta)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (32 * 24))()
rc = lib.mu_debug_bwd_trace(buf, 32 * 24)
assert rc == 0, rc
t = torch.tensor(list(buf), dtype=torch.int64).view(32, 24)
base = t[:, 8]
print("period (s_full to s_full):", (base[1:] - base[:-1]).tolist())
order = [0, 1, 7, 8, 19, 9, 10, 20, 15, 16, 17, 18, 11, 21, 2, 3, 4, 5, 6, 12, 13, 14]
print("tile " + " ".join(f"{NAMES[e][:11]:>11s}" for e in order))
for i in range(4, 20):
    print(f"{i + 8:4d} " + " ".join(f"{int(t[i, e] - base[i]):11d}" for e in order))
