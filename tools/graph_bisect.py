#!/usr/bin/env python
"""Which stage of the train step invalidates a CUDA-graph capture?  (debug helper for Trainer(cuda_graph=True))"""
import os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskunet_b200
from maskunet_b200 import ops
from maskunet_b200.train import Trainer

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = maskunet_b200.UNet(3, 19, compute_dtype=torch.bfloat16, channels_last=True).to(dev).to(memory_format=torch.channels_last)
tr = Trainer(net, cuda_graph=False)
tr.optimizer = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, fused=True, capturable=True)
x = torch.rand(2, 3, 128, 128, device=dev)
y = torch.randint(0, 19, (2, 128, 128), device=dev)
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()


def alive(tag):
    try:
        c = torch.cuda.is_current_stream_capturing()
        print(f"  ok after {tag} (capturing={c})", flush=True)
        return True
    except Exception as e:
        print(f"  INVALIDATED after {tag}: {str(e).splitlines()[0]}", flush=True)
        return False


def attempt(name, body):
    print("attempt:", name, flush=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.stream(s):
            g.capture_begin()
            try:
                body()
            finally:
                try:
                    g.capture_end()
                    print("  capture_end ok", flush=True)
                except Exception as e:
                    print("  capture_end failed:", str(e).splitlines()[0], flush=True)
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()


hooks = []
def add_hooks():
    for n, m in net.named_modules():
        if len(list(m.children())) == 0 or isinstance(m, maskunet_b200.Mask2FormerAttention):
            hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: alive("fwd " + n)))


def fwd_only():
    net.zero_grad(set_to_none=True)
    alive("zero_grad")
    with torch.no_grad():
        net(x)
    alive("forward(no_grad)")

def fwd_bwd():
    tr.optimizer.zero_grad(set_to_none=True)
    out = net(x)
    alive("forward")
    with torch.no_grad():
        loss, dpad = ops.cross_entropy_fused(net._padded_logits.detach(), y, -100, 19)
    alive("ce")
    net._padded_logits.backward(dpad)
    alive("backward")

def full():
    fwd_bwd()
    tr.optimizer.step()
    alive("adamw")

attempt("forward only, no_grad", fwd_only)
attempt("forward + backward", fwd_bwd)
attempt("full step", full)
if "--modules" in sys.argv:
    add_hooks()
    attempt("forward with per-module checks", fwd_only)
