#!/usr/bin/env python
"""Projection kernels alone at the 16384-token site (B = 256, C = 64): ms and GB/s of the forward projection."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops  # noqa: E402
from tools.bench_kernels import time_fn  # noqa: E402
dev = torch.device("cuda", 0)
for B, N, C in ((256, 16384, 64), (256, 4096, 64), (256, 4096, 128), (256, 1024, 256), (256, 256, 256)):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, N, C, device=dev, generator=g).bfloat16()
    w = torch.randn(3 * C, C, device=dev, generator=g) * 0.1
    b = torch.randn(3 * C, device=dev, generator=g)
    bits = (torch.rand(B, N, device=dev, generator=g) < 0.5).to(torch.int64)
    _, n_keep, keep_idx, keep_rank = ops.mask_binarize(bits)
    ms = time_fn(lambda: ops.qkv_project(x, w, b, keep_rank, n_keep, True), iters=10, warm=3)
    nbytes = 2 * B * N * C * 2 + 2 * float(n_keep.sum()) * C * 2
    # backward: dx = dz + [dq|dk|dv] W (P2), dW (P3), db -- one C-ABI call, three kernels; token-space dq / dk / dv
    dz, dq, dk, dv = (torch.randn(B, N, C, device=dev, generator=g).bfloat16() for _ in range(4))
    ms_b = time_fn(lambda: ops.qkv_project_bwd(x, dz, dq, dk, dv, w, True), iters=10, warm=3)
    nbytes_b = (4 + 1 + 4) * B * N * C * 2          # P2 reads dz, dq, dk, dv and writes dx; P3 + db read dq, dk, dv, x
    print(json.dumps({"B": B, "N": N, "C": C, "qkv_project_ms": round(ms, 4), "GB/s": round(nbytes / ms / 1e6, 1),
                      "qkv_project_bwd_ms": round(ms_b, 4), "bwd_GB/s": round(nbytes_b / ms_b / 1e6, 1)}), flush=True)
