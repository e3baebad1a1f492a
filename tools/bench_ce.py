#!/usr/bin/env python
"""Fused cross-entropy (A14) on the class-padded logits of the benchmark step: GB/s against the HBM peak.

    python tools/bench_ce.py [--batch 256]      (MU_CE_LPR=4|8: lanes per row, A/B)
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops  # noqa: E402
from tools.bench_kernels import time_fn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
dev = torch.device("cuda", 0)
for C, P in ((150, 160), (133, 160), (19, 32)):
    B, H, W = a.batch, 128, 128
    logits = (torch.randn(B, H, W, P, device=dev) * 3).bfloat16().permute(0, 3, 1, 2)
    labels = torch.randint(0, C, (B, H, W), device=dev)
    count = torch.full((1,), float(B * H * W), device=dev)
    ms = time_fn(lambda: ops.cross_entropy_fused(logits, labels, -100, C, count))
    gb = (2 * B * H * W * P * 2 + B * H * W * 8) / 1e9
    print(json.dumps({"kernel": "cross_entropy_fused", "B": B, "classes": C, "pitch": P, "ms": round(ms, 4),
                      "GBs": round(gb / ms * 1e3, 1), "frac_of_6537": round(gb / ms * 1e3 / 6537, 3)}), flush=True)
