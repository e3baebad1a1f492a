#!/usr/bin/env python
"""Is a kernel power-bound?  Loops one kernel for a few seconds while a thread samples NVML (SM clock, power draw)
every few milliseconds; prints the time of the first launches after idle (burst clock) against the sustained ones.

    python tools/power_probe.py [--seconds 3] [--batch 64] [--what fwd|bwd|gemm|all]
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskunet_b200 import ops  # noqa: E402
from tools.bench_kernels import make  # noqa: E402
import pynvml  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self, period=0.004):
        super().__init__(daemon=True)
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.period, self.rows, self.stop_flag = period, [], False

    def run(self):
        while not self.stop_flag:
            try:
                self.rows.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM),
                                  pynvml.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except pynvml.NVMLError:
                pass
            time.sleep(self.period)


def probe(name, fn, flops, seconds):
    torch.cuda.synchronize()
    time.sleep(1.0)                                    # idle: clocks recover
    s = Sampler()
    s.start()
    evs = []
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
    s.stop_flag = True
    s.join()
    ms = [a.elapsed_time(b) for a, b in evs]
    n = len(ms)
    first, last = ms[:4], ms[n // 2:]
    clk = [r[1] for r in s.rows]
    pw = [r[2] for r in s.rows]
    half = len(clk) // 2
    rec = {"kernel": name, "launches": n, "first_ms": [round(x, 4) for x in first],
           "sustained_ms_median": round(sorted(last)[len(last) // 2], 4),
           "tflops_first": round(flops / (min(first) * 1e-3) / 1e12, 1),
           "tflops_sustained": round(flops / (sorted(last)[len(last) // 2] * 1e-3) / 1e12, 1),
           "sm_mhz_first_samples": clk[:6], "sm_mhz_sustained_median": sorted(clk[half:])[len(clk[half:]) // 2] if clk else None,
           "power_w_max": round(max(pw), 1) if pw else None,
           "power_w_sustained_median": round(sorted(pw[half:])[len(pw[half:]) // 2], 1) if pw else None}
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--what", default="all")
    a = ap.parse_args()
    B, N, C = a.batch, 16384, 64
    q, kc, vc, n_keep, keep_idx = make(B, N, C)
    nk = float(n_keep.sum())
    if a.what in ("fwd", "all"):
        probe("attn_fwd", lambda: ops.attn_fwd(q, kc, vc, n_keep), 4 * N * nk * C, a.seconds)
    if a.what in ("bwd", "all"):
        o, lse = ops.attn_fwd(q, kc, vc, n_keep)
        d_o = torch.randn_like(o)
        delta = (d_o.float() * o.float()).sum(-1)
        probe("attn_bwd", lambda: ops.attn_bwd(q, kc, vc, n_keep, keep_idx, d_o, lse, delta), 8 * N * nk * C, a.seconds)
    if a.what in ("gemm", "all"):
        x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        y = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
        probe("cublas_bf16_8192", lambda: torch.matmul(x, y), 2 * 8192 ** 3, a.seconds)


if __name__ == "__main__":
    main()
