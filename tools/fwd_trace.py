#!/usr/bin/env python
"""Timeline of ONE attention-forward CTA (trace build of the library, -DMU_FWD_TRACE=1):

    python maskunet_b200/build.py --variant maskunet_b200/build_variant_trace.so MU_FWD_TRACE=1
    MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_trace.so python tools/fwd_trace.py

SM-clock offsets of the pipeline events of CTA (0, 0), per key tile, relative to the tile's 'softmax: s_full' event."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops, _lib  # noqa: E402
from tools.bench_kernels import make  # noqa: E402

NAMES = ["mma:K landed", "mma:S issued", "mma:p_full", "mma:PV issued", "sm:wait s", "sm:s_full", "sm:in regs",
         "sm:max done", "sm:32 exps", "sm:P free", "sm:exps done", "tma:K issue", "tma:V issue", "mma:V landed",
         "sm:published", "-"]
B, N, C = 16, 16384, 64
q, kc, vc, n_keep, keep_idx = make(B, N, C)
for _ in range(3):
    ops.attn_fwd(q, kc, vc, n_keep)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (32 * 16))()
assert lib.mu_debug_fwd_trace(buf, 32 * 16) == 0
t = torch.tensor(list(buf), dtype=torch.int64).view(32, 16)
base = t[:, 5]
print("period (s_full to s_full):", (base[1:] - base[:-1]).tolist())
order = [11, 12, 0, 1, 4, 5, 6, 7, 8, 9, 10, 14, 13, 2, 3]
print("tile " + " ".join(f"{NAMES[e][:12]:>12s}" for e in order))
for i in range(4, 20):
    print(f"{i + 8:4d} " + " ".join(f"{int(t[i, e] - base[i]):12d}" for e in order))
