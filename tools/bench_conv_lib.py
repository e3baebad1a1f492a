"""Library (cuDNN) timings of every conv3x3 layer shape of UNet(3,150) at batch 256, bf16 channels-last:
the target our own K7 kernels are measured against.  Prints one JSON line per shape."""
import json, sys, torch
import torch.nn.functional as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SHAPES = [(64, 64, 128), (128, 128, 128), (128, 64, 128), (64, 64, 64), (64, 128, 64), (128, 128, 64),
          (256, 256, 64), (256, 128, 64), (128, 64, 64), (128, 128, 32), (128, 256, 32), (256, 256, 32),
          (512, 512, 32), (512, 256, 32), (256, 128, 32), (256, 256, 16), (256, 512, 16), (512, 512, 16),
          (512, 256, 16)]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0.record()
    for _ in range(n):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n


torch.backends.cudnn.benchmark = True
for cin, cout, hw in (SHAPES if __name__ == "__main__" else []):
    x = torch.randn(B, cin, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(cout, cin, 3, 3, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True); w.requires_grad_(True)
    y = F.conv2d(x, w, padding=1)
    dy = torch.randn_like(y)
    fl = 2.0 * B * hw * hw * cin * cout * 9
    tf = timeit(lambda: F.conv2d(x, w, padding=1))
    td = timeit(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
    tw = timeit(lambda: torch.autograd.grad(y, w, dy, retain_graph=True))
    print(json.dumps({"cin": cin, "cout": cout, "hw": hw, "fwd_ms": round(tf, 3), "dgrad_ms": round(td, 3),
                      "wgrad_ms": round(tw, 3), "fwd_tf": round(fl / tf / 1e9, 1), "dgrad_tf": round(fl / td / 1e9, 1),
                      "wgrad_tf": round(fl / tw / 1e9, 1)}), flush=True)
    del x, w, y, dy
