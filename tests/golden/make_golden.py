#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It AST-loads the reference's classes (oracle/ref_loader.py), runs them on
seeded inputs and stores inputs + outputs.  The reference tree does not travel
to the GPU box, these files do.  Seeds and recipes follow SURVEY.md 8(c).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_loader import load_reference_classes  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

ATTN_CASES = [  # name, B, C, H, W
    ("attn_b2_c64_8x8", 2, 64, 8, 8),
    ("attn_b2_c128_16x16", 2, 128, 16, 16),
    ("attn_b1_c256_16x16", 1, 256, 16, 16),
    ("attn_b3_c64_16x8", 3, 64, 16, 8),       # ragged: H != W, N = 128 = exactly one tile
    ("attn_b2_c64_20x20", 2, 64, 20, 20),     # N = 400: not a multiple of the 128-row tile
]


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def make_attention(ref):
    for name, B, C, H, W in ATTN_CASES:
        torch.manual_seed(1234)
        m = ref.Mask2FormerAttention(C, C)
        x = torch.randn(B, C, H, W, requires_grad=True)
        y = m(x)                                   # draws the mask (ade_semantic.py:178)
        keep = (m.mask[:, 0, :] == 0)
        # KAT scalars as recorded in SURVEY.md 8(c): loss = y.square().sum()
        (gx_sq,) = torch.autograd.grad(y.square().sum(), x, retain_graph=True)
        m.zero_grad()
        y.square().sum().backward(retain_graph=True)
        dwq_sq = m.query.weight.grad.abs().sum().item()
        m.zero_grad()
        x.grad = None
        # well-conditioned gradient check: loss = <y, dy> with a fixed random dy
        gen = torch.Generator().manual_seed(4321)
        dy = torch.randn(y.shape, generator=gen)
        (y * dy).sum().backward()
        out = dict(x=np32(x), y=np32(y), dy=np32(dy), dx=np32(x.grad),
                   keep=keep.numpy().astype(np.uint8),
                   kat=np.array([y.abs().sum().item(), gx_sq.abs().sum().item(), dwq_sq], dtype=np.float64))
        for k, p in m.named_parameters():
            out[f"param.{k}"] = np32(p)
            out[f"grad.{k}"] = np32(p.grad)
        np.savez(os.path.join(HERE, name + ".npz"), **out)
        print(name, "keep", keep.sum(1).tolist(), "y.abs.sum", out["kat"][0])


def param_digest(sd):
    return {k: float(v.double().abs().sum()) for k, v in sd.items() if v.dtype.is_floating_point}


def make_unet(script, variant, c_out, fname, batch=2):
    ref = load_reference_classes(script)
    torch.manual_seed(1234)
    model = ref.UNet(3, c_out)
    digest = param_digest(model.state_dict())
    model.eval()
    x = torch.rand(batch, 3, 128, 128)
    with torch.no_grad():
        outs = model(x)
    outs = outs if isinstance(outs, tuple) else (outs,)
    keeps = {n: (getattr(model, n).mask[:, 0, :] == 0) for n, _, _ in uo.ATTN_SITES}
    gen = torch.Generator().manual_seed(99)
    store = {"x_seed_note": np.array([1234], dtype=np.int64)}
    for i, o in enumerate(outs):
        flat = o.reshape(-1)
        idx = torch.randint(0, flat.numel(), (8192,), generator=gen)
        store[f"out{i}.shape"] = np.array(o.shape, dtype=np.int64)
        store[f"out{i}.sample_idx"] = idx.numpy()
        store[f"out{i}.sample"] = np32(flat[idx])
        store[f"out{i}.stats"] = np.array([o.double().sum().item(), o.abs().max().item(),
                                          (o == 0).double().mean().item()], dtype=np.float64)
    store["argmax"] = outs[0].argmax(dim=1).numpy().astype(np.uint8)
    top2 = outs[0].topk(2, dim=1).values
    store["argmax_margin"] = np32(top2[:, 0] - top2[:, 1]).astype(np.float16)
    for n, k in keeps.items():
        store[f"keep.{n}"] = np.packbits(k.numpy().astype(np.uint8), axis=1)

    # ---- gradient goldens.  Every phase starts from the initial state (fresh BatchNorm running statistics), with the
    # masks drawn above, dropout disabled, labels from seed 1.
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    labels = torch.randint(0, c_out, (batch, 128, 128), generator=torch.Generator().manual_seed(1))

    def step(mode, rounded=False, autocast=False):
        """loss, logits, {name: grad} of one forward + backward of the reference's own classes.
        mode 'train': batch-statistics BatchNorm (the benchmarked step); 'evalgrad': model.eval() with autograd on
        (frozen BatchNorm).  rounded: parameters and input rounded to bf16, arithmetic still fp32 -- the perturbation
        ANY bf16 implementation applies before it computes anything.  autocast: the reference under
        torch.autocast('cpu', bfloat16), i.e. the reference's own bf16 arithmetic."""
        model.load_state_dict(sd0)
        model.zero_grad(set_to_none=True)
        model.train(mode == "train")
        model.dropout.p = 0.0
        xi = x
        if rounded:
            xi = x.bfloat16().float()
            with torch.no_grad():
                for p in model.parameters():
                    p.copy_(p.bfloat16().float())
        if autocast:
            with torch.autocast("cpu", dtype=torch.bfloat16):
                o = model(xi)
        else:
            o = model(xi)
        o = tuple(t.float() for t in o) if isinstance(o, tuple) else (o.float(),)
        loss = torch.nn.functional.cross_entropy(o[0], labels)
        if len(o) == 3:     # instance variant: every head gets a gradient
            loss = loss + 0.5 * o[1].square().mean() + 0.5 * o[2].square().mean()
        loss.backward()
        grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
        return float(loss), o[0].detach(), grads

    def vec_err(ga, gb):
        ks = [k for k in ga if ga[k] is not None]
        num = torch.cat([(ga[k] - gb[k]).flatten() for k in ks]).double().norm()
        return float(num / torch.cat([ga[k].flatten() for k in ks]).double().norm())

    def norm_errs(ga, gb):
        out = []
        for k in ga:
            if ga[k] is None or k.endswith("key.bias"):
                continue
            na, nb = float(ga[k].double().norm()), float(gb[k].double().norm())
            if na > 1e-7:
                out.append(abs(nb - na) / na)
        return np.array(out)

    sgen = torch.Generator().manual_seed(7)
    names = [k for k, _ in model.named_parameters()]
    for mode in ("train", "evalgrad"):
        loss, logits, grads = step(mode)
        norms = [float(grads[k].double().norm()) if grads[k] is not None else -1.0 for k in names]
        store[f"{mode}.loss"] = np.array([loss], dtype=np.float64)
        store[f"{mode}.grad_norms"] = np.array(norms, dtype=np.float64)
        store[f"{mode}.logits_stats"] = np.array([logits.double().sum().item(), logits.abs().max().item()], dtype=np.float64)
        flat = logits.reshape(-1)
        lidx = torch.randint(0, flat.numel(), (8192,), generator=sgen)
        store[f"{mode}.logits_sample_idx"] = lidx.numpy()
        store[f"{mode}.logits_sample"] = np32(flat[lidx])
        # a fixed 256-element sample of every parameter gradient (enough for a whole-vector error estimate)
        idx_of = {}
        for k in names:
            if grads[k] is None:
                continue
            n = grads[k].numel()
            idx_of[k] = torch.randint(0, n, (256,), generator=sgen) if n > 256 else torch.arange(n)

        def sample(g):
            return torch.cat([g[k].reshape(-1)[i] for k, i in idx_of.items()])

        store[f"{mode}.grad_sample_idx"] = np.concatenate([i.numpy().astype(np.int64) for i in idx_of.values()])
        store[f"{mode}.grad_sample"] = np32(sample(grads))
        # the reference's OWN sensitivity to bf16: what a bf16 implementation can be held to at network level
        for tag, kw in (("rounded", dict(rounded=True)), ("autocast", dict(autocast=True))):
            l2, lg2, g2 = step(mode, **kw)
            ne = norm_errs(grads, g2)
            sv = float((sample(g2) - sample(grads)).double().norm() / sample(grads).double().norm())
            store[f"{mode}.sens.{tag}"] = np.array([
                abs(l2 - loss) / abs(loss), float((lg2 - logits).norm() / logits.norm()), vec_err(grads, g2),
                float(np.median(ne)), float(np.quantile(ne, 0.9)), float(ne.max()), sv], dtype=np.float64)
            print(fname, mode, tag, "loss/logits/gradvec/norm-median/norm-p90/norm-max/sample-vec", store[f"{mode}.sens.{tag}"])
    model.load_state_dict(sd0)
    loss = store["train.loss"][0]
    np.savez_compressed(os.path.join(HERE, fname + ".npz"), **store)
    with open(os.path.join(HERE, fname + ".json"), "w") as fh:
        json.dump({"script": script, "variant": variant, "c_out": c_out, "seed": 1234, "batch": batch,
                   "torch": torch.__version__, "param_digest": digest, "grad_names": names}, fh, indent=0)
    print(fname, "out0 stats", store["out0.stats"], "loss", float(loss))


def make_postproc():
    """mean_iou and the class map of the reference's own function (ade_semantic.py:128-146) on seeded logits:
    ReLU'd logits (exact zeros and ties, as the head produces), an ignore label, a class that never occurs."""
    from oracle.ref_loader import load_reference_functions
    fn = load_reference_functions("ade_semantic", ("mean_iou",)).mean_iou
    for name, B, C, H, W, seed in (("postproc_c150", 2, 150, 16, 16, 7), ("postproc_c19", 3, 19, 8, 32, 8)):
        g = torch.Generator().manual_seed(seed)
        logits = torch.relu(torch.randn(B, C, H, W, generator=g)).to(torch.bfloat16).float()   # bf16-representable
        labels = torch.randint(0, C - 1, (B, H, W), generator=g)                               # class C-1 never labelled
        labels[0, 0, :5] = 255
        miou = fn(logits, labels, C)
        pred = torch.argmax(torch.softmax(logits / 0.5, dim=1), dim=1)                          # :130-131 verbatim recipe
        np.savez_compressed(os.path.join(HERE, name + ".npz"), logits=np32(logits), labels=labels.numpy(),
                            pred=pred.numpy().astype(np.int16), miou=np.array([float(miou)], dtype=np.float64))
        print(name, float(miou))


INSTANCE_LOSS_CASES = [  # name, script, ignore_value, B, C, H, W, seed
    ("instance_loss_coco", "coco_panoptic", None, 4, 13, 16, 16, 11),
    ("instance_loss_city", "city_instance", 255, 3, 19, 8, 24, 12),
    ("instance_loss_blobs", "coco_panoptic", None, 8, 7, 16, 32, 13),
]


def instance_mask_case(B, H, W, gen, blobs):
    """Ids as the datasets produce them: 0 background, large panoptic segment ids, 255, a one-pixel instance."""
    if blobs:                                    # rectangles: the first two pixels of an instance are neighbours
        im = torch.zeros(B, H, W, dtype=torch.int64)
        for b in range(B):
            for j in range(3):
                h0, w0 = int(torch.randint(0, H - 4, (1,), generator=gen)), int(torch.randint(0, W - 6, (1,), generator=gen))
                im[b, h0:h0 + 4, w0:w0 + 6] = 1_000_003 * (b + 1) + j
    else:
        im = torch.randint(0, 7, (B, H, W), generator=gen) * 4099
        im[im == 2 * 4099] = 0
    im[0, 0, :3] = 255
    im[B - 1, H - 1, W - 1] = 77_777_777         # fewer than two pixels: skipped (:498)
    return im


def make_instance_loss():
    """InstanceContrastiveLoss of the reference's own classes (coco_panoptic.py:482-521, city_instance.py:279-307):
    loss, gradient and the state of the CPU generator after the call (it draws one randint per instance, :510)."""
    for name, script, ignore, B, C, H, W, seed in INSTANCE_LOSS_CASES:
        Ref = load_reference_classes(script, ("InstanceContrastiveLoss",)).InstanceContrastiveLoss
        gen = torch.Generator().manual_seed(seed)
        amp = 0.0625 if "blobs" in name else 1.0                  # small logits: the hinge stays active with d(a, p) ~ 0
        sem = (amp * torch.relu(torch.randn(B, C, H, W, generator=gen))).to(torch.bfloat16).float().requires_grad_()
        im = instance_mask_case(B, H, W, gen, "blobs" in name)
        torch.manual_seed(1000 + seed)
        loss = Ref()(sem, im)
        next_draw = int(torch.randint(0, 2 ** 31, (1,)))          # fingerprint of the generator state after the call
        loss.backward()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sem=np32(sem), instance_mask=im.numpy(),
                            loss=np.array([float(loss)], dtype=np.float64), grad=np32(sem.grad),
                            meta=np.array([1000 + seed, -1 if ignore is None else ignore, next_draw], dtype=np.int64))
        print(name, float(loss), "grad.abs.sum", float(sem.grad.abs().sum()))


def make_to_tensor():
    """torchvision's own ToTensor (the reference's transform, ade_semantic.py:9,85) on every uint8 value and on a
    random RGB image."""
    from torchvision.transforms import ToTensor
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, size=(12, 20, 3), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "to_tensor.npz"), ramp=ramp, ramp_out=ToTensor()(ramp).numpy(),
                        img=img, img_out=ToTensor()(img).numpy())
    print("to_tensor", float(ToTensor()(img).sum()))


RESIZE_CASES = [  # name, source (H, W), output (H, W): down-scaling, the exact-2x INTER_AREA route, enlarging, mixed
    ("ade_like", (150, 233), (128, 128)), ("exact_2x", (256, 256), (128, 128)), ("enlarge", (37, 50), (128, 128)),
    ("wide", (96, 512), (128, 128)), ("tall_to_rect", (75, 100), (64, 96)),
]


def make_resize():
    """cv2's OWN outputs for the two resize calls of the dataset classes (ade_semantic.py:72-73) and ToTensor on the
    resized image (:75-76): INTER_LINEAR on an RGB uint8 image, INTER_NEAREST on a uint8 label map."""
    import cv2
    from torchvision.transforms import ToTensor
    rng = np.random.default_rng(21)
    store = {"cv2_version": np.array([int(v) for v in cv2.__version__.split(".")[:3]], dtype=np.int64)}
    for name, (sh, sw), (oh, ow) in RESIZE_CASES:
        img = rng.integers(0, 256, size=(sh, sw, 3), dtype=np.uint8)
        img[: sh // 2] = (np.linspace(0, 255, sw)[None, :, None] * np.ones((sh // 2, 1, 3))).astype(np.uint8)   # ramps too
        mask = rng.integers(0, 151, size=(sh, sw), dtype=np.uint8)
        lin = cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)
        near = cv2.resize(mask, (ow, oh), interpolation=cv2.INTER_NEAREST)
        store.update({f"{name}.img": img, f"{name}.mask": mask, f"{name}.linear": lin, f"{name}.nearest": near,
                      f"{name}.tensor": ToTensor()(lin).numpy()})
        print("resize", name, img.shape, "->", lin.shape, int(lin.sum()), int(near.sum()))
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **store)


def main():
    if "--resize-only" in sys.argv:
        make_resize()
        return
    if "--to-tensor-only" in sys.argv:
        make_to_tensor()
        return
    if "--postproc-only" in sys.argv:
        make_postproc()
        return
    if "--instance-loss-only" in sys.argv:
        make_instance_loss()
        return
    if "--unet-only" in sys.argv:
        make_unet("ade_semantic", "semantic", 150, "unet_semantic")
        make_unet("city_instance", "instance", 19, "unet_instance")
        return
    make_instance_loss()
    make_to_tensor()
    make_resize()
    make_postproc()
    ref = load_reference_classes("ade_semantic")
    make_attention(ref)
    make_unet("ade_semantic", "semantic", 150, "unet_semantic")
    make_unet("city_instance", "instance", 19, "unet_instance")


if __name__ == "__main__":
    main()
