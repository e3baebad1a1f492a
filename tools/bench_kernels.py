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


def run(B, N, C, bwd, peak, verbose=True):
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
    if verbose:
        print(json.dumps(rec), flush=True)
    return rec


def run_generalised(B, Q, N, heads, peak, C=256, verbose=True):
    """Generalised mode: K13 bits + masked multi-head cross attention, fwd and bwd.
    FLOPs: executed = every (query, key) pair of the tiles; useful = pairs whose bit is set."""
    from maskunet_b200 import query_attention as qa
    g = torch.Generator(device=DEV).manual_seed(0)
    mk = lambda *s, amp=1.0: (amp * torch.randn(*s, device=DEV, generator=g)).bfloat16()
    qe, feat = mk(B, Q, C, amp=0.25), mk(B, N, C, amp=0.25)
    q, k, v = mk(B, Q, C), mk(B, N, C), mk(B, N, C)
    d = C // heads
    ms_bits = time_fn(lambda: qa.query_mask_bits(qe, feat))
    bits, bits_t, row_count = qa.query_mask_bits(qe, feat)
    kept = float(row_count.sum())
    NKP = ops.nkp_of(N)
    qh, kh, vh = qa._to_heads(q, heads, Q), qa._to_heads(k, heads, NKP), qa._to_heads(v, heads, NKP)
    scale = d ** -0.5
    ms_f = time_fn(lambda: qa.query_attn_fwd(qh, kh, vh, bits, heads, N, scale))
    o, lse = qa.query_attn_fwd(qh, kh, vh, bits, heads, N, scale)
    d_o = torch.randn_like(o)
    delta = (d_o.float() * o.float()).sum(-1)
    ms_b = time_fn(lambda: qa.query_attn_bwd(qh, kh, vh, bits_t, d_o, lse, delta, heads, N, scale), iters=5, warm=2)
    pairs = B * Q * N
    rec = {"mode": "generalised", "B": B, "Q": Q, "N": N, "heads": heads, "head_dim": d, "kept_frac": round(kept / pairs, 3),
           "bits_ms": round(ms_bits, 4), "bits_tflops": round(2 * pairs * C / ms_bits / 1e9, 1),
           "fwd_ms": round(ms_f, 4), "fwd_tflops_useful": round(4 * kept * heads * d / ms_f / 1e9, 1),
           "fwd_tflops_executed": round(4 * pairs * heads * 64 / ms_f / 1e9, 1),
           "bwd_ms": round(ms_b, 4), "bwd_tflops_useful": round(8 * kept * heads * d / ms_b / 1e9, 1),
           "bwd_tflops_executed": round(10 * pairs * heads * 64 / ms_b / 1e9, 1), "peak_tflops": peak}
    if verbose:
        print(json.dumps(rec), flush=True)
    return rec


def run_eager(B, N, C, dtype, peak):
    """Secondary baseline of SURVEY 8(d): the reference module's attention arithmetic (ade_semantic.py:174-186) as eager
    PyTorch on this GPU -- scores, division, additive 0 / -inf mask, softmax, PV, all materialised -- fwd + bwd."""
    g = torch.Generator(device=DEV).manual_seed(0)
    q, k, v = (torch.randn(B, N, C, device=DEV, generator=g, dtype=dtype).requires_grad_() for _ in range(3))
    bias = torch.where(torch.randint(0, 2, (B, N), device=DEV, generator=g) > 0.5, 0.0, float("-inf")).to(dtype)
    mask = bias.unsqueeze(1).expand(-1, N, -1)
    nk = float((bias == 0).sum())

    def fwd():
        s = torch.matmul(q, k.transpose(-2, -1)) / (C ** 0.5)
        return torch.matmul(torch.softmax(s + mask, dim=-1), v)

    def fwd_bwd():
        out = fwd()
        out.backward(torch.ones_like(out))
        q.grad = k.grad = v.grad = None

    with torch.no_grad():
        ms_f = time_fn(fwd, iters=5, warm=2)
    ms_fb = time_fn(fwd_bwd, iters=5, warm=2)
    rec = {"mode": "eager reference arithmetic", "dtype": str(dtype).replace("torch.", ""), "B": B, "N": N, "C": C,
           "fwd_ms": round(ms_f, 3), "fwd_tflops_useful": round(4 * N * nk * C / ms_f / 1e9, 1),
           "fwd_bwd_ms": round(ms_fb, 3), "fwd_bwd_tflops_useful": round(12 * N * nk * C / ms_fb / 1e9, 1),
           "peak_tflops": peak}
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--out", default=None)
    ap.add_argument("--site", type=int, default=None, help="run only this entry of the site list (for ncu)")
    ap.add_argument("--generalised", action="store_true", help="only the generalised-mode sweep (heads, queries)")
    ap.add_argument("--eager", action="store_true", help="only the eager-PyTorch reference arithmetic on this GPU")
    ap.add_argument("--deterministic", action="store_true", help="fixed-order dQ accumulation (maskunet_b200.set_deterministic)")
    args = ap.parse_args()
    if args.deterministic:
        import maskunet_b200
        maskunet_b200.set_deterministic(True)
    peak = 1601.0
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peak = json.load(open(pk))["bf16_tflops"]
    recs = []
    if args.generalised or args.eager:
        if args.generalised:
            for heads in (4, 8):
                for Q in (50, 100, 200):
                    for N in (1024, 4096):
                        recs.append(run_generalised(64, Q, N, heads, peak))
            recs.append(run_generalised(16, 200, 16384, 4, peak))
        if args.eager:
            for N, C, B in ((16384, 64, 4), (4096, 128, 32), (1024, 256, 128)):
                for dt in (torch.float32, torch.bfloat16):
                    recs.append(run_eager(B, N, C, dt, peak))
        if args.out:
            json.dump({"peak_tflops_burst": peak, "records": recs}, open(args.out, "w"), indent=1)
        return
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
