"""Our K7 kernels on every conv3x3 layer shape of UNet(3,150) at batch 256 (same list as bench_conv_lib.py)."""
import json, sys, torch
sys.path.insert(0, ".")
from maskunet_b200 import ops
from tools.bench_conv_lib import SHAPES, timeit  # noqa

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if __name__ == "__main__":
    for cin, cout, hw in SHAPES:
        x = torch.randn(B, cin, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = torch.randn(cout, cin, 3, 3, device="cuda")
        dy = torch.randn(B, cout, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wf, wd = ops.conv_prep_weights(w, True)
        fl = 2.0 * B * hw * hw * cin * cout * 9
        tf = timeit(lambda: ops.conv3x3_fwd(x, wf, True))
        td = timeit(lambda: ops.conv3x3_bwd_data(dy, wd))
        tw = timeit(lambda: ops.conv3x3_bwd_weight(x, dy))
        print(json.dumps({"cin": cin, "cout": cout, "hw": hw, "fwd_ms": round(tf, 3), "dgrad_ms": round(td, 3),
                          "wgrad_ms": round(tw, 3), "fwd_tf": round(fl / tf / 1e9, 1),
                          "dgrad_tf": round(fl / td / 1e9, 1), "wgrad_tf": round(fl / tw / 1e9, 1)}), flush=True)
        del x, w, dy
