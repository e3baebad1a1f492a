"""HBM-bound kernels at the largest activation shape of the step ([256, C, 128, 128] bf16): achieved GB/s vs 6537."""
import json, sys, torch
sys.path.insert(0, ".")
from maskunet_b200 import ops
from tools.bench_conv_lib import timeit

B = 256
for C, HW in ((64, 128), (128, 128), (256, 64), (160, 128)):
    x = torch.randn(B, C, HW, HW, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    r = torch.randn_like(x)
    dy = torch.randn_like(x)
    g = torch.rand(C, device="cuda") + 0.5
    b = torch.randn(C, device="cuda")
    nbytes = x.numel() * 2
    sums = ops.column_sums(x)
    rec = {"C": C, "HW": HW, "GB": round(nbytes / 1e9, 3)}
    for act, name in ((ops.ACT_GELU, "gelu"), (ops.ACT_NONE, "none"), (ops.ACT_RELU, "relu")):
        for res in (None, r):
            tag = name + ("+res" if res is not None else "")
            y, mean, rstd, a, bb = ops.bn_act_fwd_stats(x, res, g, b, sums, 1e-5, act)
            t = timeit(lambda: ops.bn_act_fwd_stats(x, res, g, b, sums, 1e-5, act))
            passes = 2 + (1 if res is not None else 0)
            rec["fwd_" + tag] = round(passes * nbytes / t / 1e6, 0)
            t = timeit(lambda: ops.bn_act_bwd(dy, x, res, a, bb, mean, rstd, act))
            passes = 5 + (2 if res is not None else 0)
            rec["bwd_" + tag] = round(passes * nbytes / t / 1e6, 0)
    t = timeit(lambda: ops.column_sums(x))
    rec["stats"] = round(nbytes / t / 1e6, 0)
    print(json.dumps(rec), flush=True)
