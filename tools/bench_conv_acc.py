"""Data gradient of the six residual ConvBlocks' first convolution at batch 256: plain kernel + autograd's add pass
against the accumulating kernel (TMA bf16 reduce-add into the residual branch's gradient)."""
import json, sys, torch
sys.path.insert(0, ".")
from maskunet_b200 import ops
from tools.bench_kernels import time_fn  # noqa

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SHAPES = [(64, 64, 64), (128, 128, 32), (256, 256, 16), (512, 512, 32), (256, 256, 64), (128, 128, 128)]
if __name__ == "__main__":
    for cin, cout, hw in SHAPES:
        w = torch.randn(cout, cin, 3, 3, device="cuda")
        dy = torch.randn(B, cout, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        other = torch.randn(B, cin, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        _, wd = ops.conv_prep_weights(w, True)
        t_plain = time_fn(lambda: ops.conv3x3_bwd_data(dy, wd))
        dx = ops.conv3x3_bwd_data(dy, wd)
        t_add = time_fn(lambda: torch.add(dx, other))
        t_acc = time_fn(lambda: ops.conv3x3_bwd_data_acc(dy, wd, other))
        print(json.dumps({"cin": cin, "cout": cout, "hw": hw, "dgrad_ms": round(t_plain, 4), "add_ms": round(t_add, 4),
                          "dgrad_plus_add_ms": round(t_plain + t_add, 4), "dgrad_acc_ms": round(t_acc, 4)}), flush=True)
        del w, dy, other, dx
