"""Residual LayerNorm with the fused re-view against the token-major kernels + the separate transpose pass, alone
(CUDA events, L2 flushed), at the three large site shapes of the benchmark batch."""
import json, sys, torch
sys.path.insert(0, ".")
from maskunet_b200 import ops
from tools.bench_kernels import time_fn  # noqa

dev = torch.device("cuda", 0)
for B, N, C in ((256, 16384, 64), (256, 4096, 128), (256, 1024, 256), (256, 4096, 64)):
    g = torch.Generator(device=dev).manual_seed(0)
    o, x, dy = (torch.randn(B, N, C, device=dev, generator=g).bfloat16() for _ in range(3))
    gamma = torch.rand(C, device=dev) + 0.5
    beta = torch.randn(C, device=dev)
    y, mean, rstd = ops.residual_ln_fwd(o, x, gamma, beta, 1e-5, True)
    rec = {"B": B, "N": N, "C": C}
    rec["fwd_tok_ms"] = round(time_fn(lambda: ops.residual_ln_fwd(o, x, gamma, beta, 1e-5, True)), 4)
    rec["transpose_ms"] = round(time_fn(lambda: ops.transpose(y.view(B, C, N))), 4)
    rec["fwd_view_ms"] = round(time_fn(lambda: ops.residual_ln_fwd(o, x, gamma, beta, 1e-5, True, True)), 4)
    rec["bwd_tok_ms"] = round(time_fn(lambda: ops.residual_ln_bwd(dy, o, x, mean, rstd, gamma, True)), 4)
    rec["bwd_view_ms"] = round(time_fn(lambda: ops.residual_ln_bwd(dy, o, x, mean, rstd, gamma, True, True)), 4)
    gb = B * N * C * 2 / 1e9
    rec["fwd_view_GB/s"] = round(3 * gb / rec["fwd_view_ms"] * 1e3, 0)
    rec["bwd_view_GB/s"] = round(4 * gb / rec["bwd_view_ms"] * 1e3, 0)
    print(json.dumps(rec), flush=True)
