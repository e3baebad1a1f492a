#!/usr/bin/env python
"""Kernel sweep (BASELINE.json configs[4]): mask-attention forward/backward TFLOP/s vs tensor peak.

    python tools/bench_kernels.py [--sites] [--sweep] [--bwd] [--batch B]

Useful FLOPs only (kept keys): fwd 4 N n_keep C, bwd 8 N n_keep C.  CUDA events, 3 warm-ups, L2 flushed
between timed launches by writing a 256 MB buffer.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops  # noqa: E402

DEV = torch.device("cuda", 0)


def make(B, N, C, p_keep=0.5, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    NKP = ops.nkp_of(N)
    q = torch.randn(B, N, C, device=DEV, generator=g).bfloat16()
    kc = torch.randn(B, NKP, C, device=DEV, generator=g).bfloat16()
    vc = torch.randn(B, NKP, C, device=DEV, generator=g).bfloat16()
    bits = (torch.rand(B, N, device=DEV, generator=g) < p_keep).to(torch.int64)
    _, n_keep, keep_idx, _ = ops.mask_binarize(bits)
    for b in range(B):
        kc[b, int(n_keep[b]):] = 0
        vc[b, int(n_keep[b]):] = 0
    return q, kc, vc, n_keep, keep_idx


def time_fn(fn, iters=10, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def run(B, N, C, bwd, peak):
    q, kc, vc, n_keep, keep_idx = make(B, N, C)
    nk = float(n_keep.sum())
    ms = time_fn(lambda: ops.attn_fwd(q, kc, vc, n_keep))
    tf = 4 * N * nk * C / (ms * 1e-3) / 1e12
    rec = {"B": B, "N": N, "C": C, "fwd_ms": round(ms, 4), "fwd_tflops": round(tf, 1), "fwd_frac": round(tf / peak, 3)}
    if bwd:
        o, lse = ops.attn_fwd(q, kc, vc, n_keep)
        d_o = torch.randn_like(o)
        delta = (d_o.float() * o.float()).sum(-1)
        ms = time_fn(lambda: ops.attn_bwd(q, kc, vc, n_keep, keep_idx, d_o, lse, delta), iters=5, warm=2)
        tf = 8 * N * nk * C / (ms * 1e-3) / 1e12
        rec.update({"bwd_ms": round(ms, 4), "bwd_tflops": round(tf, 1), "bwd_frac": round(tf / peak, 3)})
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default=None)
    ap.add_argument("--site", type=int, default=None, help="run only this entry of the site list (for ncu)")
    args = ap.parse_args()
    peak = 1601.0
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = json.load(open(pk))["bf16_tflops"]
    recs = []
    sites = [(16384, 64), (4096, 64), (4096, 128), (1024, 128), (1024, 256), (256, 256)]
    if args.site is not None:
        sites = [sites[args.site]]
    for N, C in sites:
        B = args.batch if N >= 4096 else args.batch * 8
        recs.append(run(B, N, C, args.bwd, peak))
    if args.sweep:
        for side in (8, 16, 32, 64):
            for C in (64, 128, 256):
                recs.append(run(max(args.batch, 148 * 2 * 128 // (side * side) + 1), side * side, C, args.bwd, peak))
    if args.out:
        json.dump({"peak_tflops_burst": peak, "records": recs}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
