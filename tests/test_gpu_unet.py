"""GPU parity of the full U-Net (attention on our kernels) against goldens produced by the reference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mask_attention_oracle as mao
from oracle import unet_oracle as uo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
# fp32 parity means fp32: no TF32 in the stock torch convolutions / matmuls around our kernels
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _load(fname):
    with open(os.path.join(GOLDEN, fname + ".json")) as fh:
        meta = json.load(fh)
    return meta, np.load(os.path.join(GOLDEN, fname + ".npz"))


def _build(cls, meta, z, **kw):
    torch.manual_seed(meta["seed"])
    net = cls(3, meta["c_out"], **kw)
    digest = {k: float(v.double().abs().sum()) for k, v in net.state_dict().items() if v.dtype.is_floating_point}
    bad = [k for k, v in meta["param_digest"].items() if abs(digest[k] - v) > 1e-6 * max(1.0, abs(v))]
    # same seed, same construction order, same torch build as the golden: a mismatch is an error, never a skip
    assert not bad, f"initial parameters differ from the golden's (torch RNG stream changed?): {bad[:3]}"
    x = torch.rand(meta["batch"], 3, 128, 128)
    net = net.to(DEV)
    for n, _, side in uo.ATTN_SITES:   # inject the reference's masks
        keep = torch.from_numpy(np.unpackbits(z[f"keep.{n}"], axis=1)[:, : side * side]).bool()
        getattr(net, n).mask = mao.expand_bias(mao.additive_bias(keep).to(DEV), side * side)
    return net, x


@pytest.mark.parametrize("fname,variant", [("unet_semantic", "semantic"), ("unet_instance", "instance")])
def test_unet_eval_fp32_matches_reference(fname, variant):
    import maskunet_b200
    meta, z = _load(fname)
    cls = maskunet_b200.UNet if variant == "semantic" else maskunet_b200.InstanceUNet
    net, x = _build(cls, meta, z)
    net.eval()
    with torch.no_grad():
        outs = net(x.to(DEV))
    outs = outs if isinstance(outs, tuple) else (outs,)
    for i, o in enumerate(outs):
        assert tuple(o.shape) == tuple(z[f"out{i}.shape"])
        got = o.reshape(-1)[torch.from_numpy(z[f"out{i}.sample_idx"]).to(DEV)].cpu()
        ref = torch.from_numpy(z[f"out{i}.sample"])
        err = float((got - ref).norm() / ref.norm())
        assert err < 1e-4, (i, err)
    am = outs[0].argmax(1).cpu().numpy().astype(np.uint8)
    safe = z["argmax_margin"].astype(np.float32) > 1e-3
    assert (am[safe] == z["argmax"][safe]).all()           # argmax maps identical away from numerical ties
    assert (am == z["argmax"]).mean() > 0.9995


def test_unet_train_step_fp32_grad_norms_match_reference():
    import maskunet_b200
    meta, z = _load("unet_semantic")
    net, x = _build(maskunet_b200.UNet, meta, z)
    net.train()
    net.dropout.p = 0.0
    labels = torch.randint(0, meta["c_out"], (meta["batch"], 128, 128), generator=torch.Generator().manual_seed(1))
    out = net(x.to(DEV))
    loss = torch.nn.functional.cross_entropy(out, labels.to(DEV))
    loss.backward()
    assert abs(float(loss) - float(z["train.loss"][0])) < 1e-4 * float(z["train.loss"][0])
    ref = z["train.grad_norms"]
    for (name, p), r in zip(net.named_parameters(), ref):
        if r < 0:
            assert p.grad is None, name                    # dead emb_layer parameters
            continue
        g = float(p.grad.double().norm())
        if name.endswith(("key.bias",)) or r < 1e-7:
            continue                                       # analytically zero gradients
        assert abs(g - r) < 2e-3 * r + 1e-7, (name, g, r)


def test_unet_bf16_logits_within_tolerance():
    import maskunet_b200
    meta, z = _load("unet_semantic")
    net, x = _build(maskunet_b200.UNet, meta, z, compute_dtype=torch.bfloat16)
    net.eval()
    with torch.no_grad():
        out = net(x.to(DEV)).float()
    got = out.reshape(-1)[torch.from_numpy(z["out0.sample_idx"]).to(DEV)].cpu()
    ref = torch.from_numpy(z["out0.sample"])
    err = float((got - ref).norm() / ref.norm())
    assert err < 2e-2, err
    am = out.argmax(1).cpu().numpy().astype(np.uint8)
    safe = z["argmax_margin"].astype(np.float32) > 0.1
    mismatch = float((am != z["argmax"]).mean())
    print(f"bf16: logits rel-err {err:.3e}, raw argmax mismatch fraction {mismatch:.4f}")
    assert (am[safe] == z["argmax"][safe]).mean() > 0.995


# ------------------------------------------------------------------ fused BN + activation kernels (K8)
@pytest.mark.parametrize("C,H,B", [(64, 16, 4), (150, 12, 2), (19, 9, 3), (512, 8, 2), (32, 20, 2)])
@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("with_res", [False, True])
def test_fused_bn_act_matches_torch_fp32(C, H, B, act, with_res):
    from maskunet_b200 import modules, ops
    torch.manual_seed(C + act)
    x = (torch.randn(B, C, H, H, device=DEV) * 2 + 0.5).contiguous(memory_format=torch.channels_last)
    r = torch.randn(B, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last) if with_res else None
    dy = torch.randn(B, C, H, H, device=DEV).contiguous(memory_format=torch.channels_last)
    bn_a, bn_b = torch.nn.BatchNorm2d(C).to(DEV), torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.normal_()
        bn_b.load_state_dict(bn_a.state_dict())
    outs = []
    for bn, use_fused in ((bn_a, True), (bn_b, False)):
        xi = x.clone().requires_grad_(True)
        ri = r.clone().requires_grad_(True) if with_res else None
        if use_fused:
            y = modules.fused_bn_act(xi, bn, act, ri)
        else:   # the reference's op order on NCHW tensors
            y = bn(xi.contiguous())
            if with_res:
                y = ri.contiguous() + y
            y = torch.nn.functional.gelu(y) if act == 1 else (torch.relu(y) if act == 2 else y)
        (y * dy).sum().backward()
        outs.append((y, xi.grad, ri.grad if with_res else None, bn.weight.grad, bn.bias.grad,
                     bn.running_mean.clone(), bn.running_var.clone()))
    for got, ref in zip(*outs):
        if ref is not None:
            err = float((got.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-20))
            assert err < 2e-5, err


def test_fused_bn_act_bf16_and_eval():
    from maskunet_b200 import modules
    torch.manual_seed(0)
    C = 128
    x = torch.randn(4, C, 16, 16, device=DEV).contiguous(memory_format=torch.channels_last)
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    ref = torch.nn.functional.gelu(bn(x.contiguous()))
    bn2 = torch.nn.BatchNorm2d(C).to(DEV)
    got = modules.fused_bn_act(x.bfloat16(), bn2, 1)
    assert got.dtype == torch.bfloat16 and got.is_contiguous(memory_format=torch.channels_last)
    assert float((got.float() - ref).norm() / ref.norm()) < 1e-2
    bn.eval(), bn2.eval()
    with torch.no_grad():
        ref_e = torch.relu(bn(x.contiguous()))
        got_e = modules.fused_bn_act(x, bn2, 2)
    assert float((got_e - ref_e).norm() / ref_e.norm()) < 2e-3   # running stats differ by the bf16 batch stats


def test_unet_fp32_channels_last_uses_fused_kernels_and_matches_reference():
    """Whole network in channels-last fp32: attention (fp32 CUDA-core path) + fused BN kernels vs the golden."""
    import maskunet_b200
    meta, z = _load("unet_semantic")
    net, x = _build(maskunet_b200.UNet, meta, z, channels_last=True)
    net = net.to(memory_format=torch.channels_last)
    net.train()
    net.dropout.p = 0.0
    labels = torch.randint(0, meta["c_out"], (meta["batch"], 128, 128), generator=torch.Generator().manual_seed(1))
    out = net(x.to(DEV))
    loss = torch.nn.functional.cross_entropy(out, labels.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - float(z["train.loss"][0])) < 1e-4 * float(z["train.loss"][0])
    for (name, p), r in zip(net.named_parameters(), z["train.grad_norms"]):
        if r < 1e-7 or name.endswith("key.bias"):
            continue
        g = float(p.grad.double().norm())
        # whole-network train-mode gradients cross 39 batch-statistics BatchNorms at batch 2: fp32 round-off
        # differences between the conv libraries (cuDNN here, oneDNN in the golden) are amplified to ~2e-3 on a few
        # BatchNorm weights (measured 1.9e-3 .. 2.1e-3 run to run, atomics order); the 1e-4 bar is held per kernel
        assert abs(g - r) < 5e-3 * r + 1e-7, (name, g, r)


# ------------------------------------------------------------------ K9 / K10 / K11 / A14 kernels vs the stock torch ops
def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
def test_maxpool2_matches_torch(dtype, tol):
    from maskunet_b200 import ops
    torch.manual_seed(0)
    x = _cl(torch.randn(3, 64, 20, 12, device=DEV).to(dtype)).requires_grad_(True)
    xr = x.detach().float().requires_grad_(True)
    dy = _cl(torch.randn(3, 64, 10, 6, device=DEV).to(dtype))
    y = ops.maxpool2(x)
    yr = torch.nn.functional.max_pool2d(xr, 2)
    assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y.float(), yr)
    y.backward(dy)
    yr.backward(dy.float())
    assert torch.equal(x.grad.float(), xr.grad)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("H,W", [(16, 16), (8, 12)])
def test_upsample_concat_matches_torch(dtype, tol, H, W):
    from maskunet_b200 import ops
    torch.manual_seed(1)
    x = _cl(torch.randn(2, 32, H, W, device=DEV).to(dtype)).requires_grad_(True)
    sk = _cl(torch.randn(2, 16, 2 * H, 2 * W, device=DEV).to(dtype)).requires_grad_(True)
    xr, skr = x.detach().float().requires_grad_(True), sk.detach().float().requires_grad_(True)
    dy = _cl(torch.randn(2, 48, 2 * H, 2 * W, device=DEV).to(dtype))
    y = ops.upsample_concat(sk, x)
    yr = torch.cat([skr, torch.nn.functional.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)], 1)
    assert float((y.float() - yr).norm() / yr.norm()) < tol
    y.backward(dy)
    yr.backward(dy.float())
    assert float((x.grad.float() - xr.grad).norm() / xr.grad.norm()) < tol
    assert float((sk.grad.float() - skr.grad).norm() / skr.grad.norm()) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("shape", [(2, 64, 64, 64, 64), (3, 128, 128, 32, 32), (2, 256, 256, 16, 16), (1, 16, 32, 1, 2),
                                   (2, 8, 8, 2, 1)],
                         ids=lambda s: "b%d_cs%d_cx%d_%dx%d" % s)
def test_upsample_concat_split_kernels_bit_identical_and_match_torch(dtype, tol, shape, monkeypatch):
    """Power-of-two shapes (every U-Net call) take the split-range kernels (skip copy and interpolation as separate
    warp-uniform ranges): bit-identical to the single-range kernels (MU_UPCAT_SPLIT=0), both within tolerance of torch."""
    from maskunet_b200 import ops
    B, Cs, Cx, H, W = shape
    torch.manual_seed(7)
    x = _cl(torch.randn(B, Cx, H, W, device=DEV).to(dtype))
    sk = _cl(torch.randn(B, Cs, 2 * H, 2 * W, device=DEV).to(dtype))
    dy = _cl(torch.randn(B, Cs + Cx, 2 * H, 2 * W, device=DEV).to(dtype))
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("MU_UPCAT_SPLIT", flag)
        xq, sq = x.clone(memory_format=torch.preserve_format).requires_grad_(True), \
            sk.clone(memory_format=torch.preserve_format).requires_grad_(True)
        y = ops.upsample_concat(sq, xq)
        y.backward(dy)
        res[flag] = (y.detach(), xq.grad, sq.grad)
    for a, b in zip(res["1"], res["0"]):
        assert torch.equal(a, b)
    xr, skr = x.float().requires_grad_(True), sk.float().requires_grad_(True)
    up = torch.nn.functional.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    yr = torch.cat([skr, up], 1)
    yr.backward(dy.float())
    y, gx, gs = res["1"]
    assert float((y.float() - yr).norm() / yr.norm()) < tol
    assert float((gx.float() - xr.grad).norm() / xr.grad.norm()) < tol
    assert torch.equal(gs.float(), skr.grad)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_sample_layernorm_matches_torch(dtype, tol):
    from maskunet_b200 import ops
    torch.manual_seed(2)
    B, C, H, W = 5, 64, 16, 24
    x = _cl((torch.randn(B, C, H, W, device=DEV) * 1.5 + 0.3).to(dtype)).requires_grad_(True)
    g = (torch.rand(C, H, W, device=DEV) + 0.5).requires_grad_(True)
    bt = torch.randn(C, H, W, device=DEV).requires_grad_(True)
    xr, gr, br = (t.detach().float().requires_grad_(True) for t in (x, g, bt))
    dy = _cl(torch.randn(B, C, H, W, device=DEV).to(dtype))
    y = ops.sample_layernorm(x, g.permute(1, 2, 0).contiguous(), bt.permute(1, 2, 0).contiguous(), 1e-5)[0]
    yr = torch.nn.functional.layer_norm(xr, [C, H, W], gr, br, 1e-5)
    assert float((y.float() - yr).norm() / yr.norm()) < tol
    y.backward(dy)
    yr.backward(dy.float())
    for a, b in ((x.grad, xr.grad), (g.grad, gr.grad), (bt.grad, br.grad)):
        assert float((a.float() - b).norm() / b.norm()) < tol


@pytest.mark.parametrize("C", [150, 19, 133])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_cross_entropy_fused_matches_torch(C, dtype, tol):
    from maskunet_b200 import ops
    torch.manual_seed(3)
    logits = _cl((torch.randn(3, C, 16, 20, device=DEV) * 3).to(dtype)).requires_grad_(True)
    lr = logits.detach().float().requires_grad_(True)
    labels = torch.randint(0, C, (3, 16, 20), device=DEV)
    labels[0, :4] = 255
    loss = ops.cross_entropy_fused(logits, labels, 255)[0]
    ref = torch.nn.functional.cross_entropy(lr, labels, ignore_index=255)
    assert abs(float(loss) - float(ref)) < tol * max(1.0, abs(float(ref)))
    (loss.squeeze() * 2.0).backward()
    (ref * 2.0).backward()
    assert float((logits.grad.float() - lr.grad).norm() / lr.grad.norm()) < tol
    assert (logits.grad[0, :, :4] == 0).all()


@pytest.mark.parametrize("dtype,C", [(torch.float32, 19), (torch.bfloat16, 19), (torch.bfloat16, 32)])
def test_cross_entropy_fused_out_of_range_label_poisons_loss(dtype, C):
    """nn.CrossEntropyLoss raises a device assert for a label outside [0, C) that is not ignore_index (e.g. the 255 void
    label of Cityscapes, city_semantic.py:341, under the default ignore_index=-100).  The fused kernel cannot raise:
    the loss and the offending rows' gradients are NaN, every other row keeps its gradient."""
    from maskunet_b200 import ops
    torch.manual_seed(5)
    logits = _cl(torch.randn(2, C, 8, 16, device=DEV).to(dtype))
    labels = torch.randint(0, C, (2, 8, 16), device=DEV)
    labels[1, 3, :5] = 255
    loss, dl = ops.cross_entropy_fused(logits, labels, -100)
    assert bool(torch.isnan(loss).all())
    assert bool(torch.isnan(dl[1, :, 3, :5]).all())
    ok = torch.ones(2, 8, 16, dtype=torch.bool, device=DEV)
    ok[1, 3, :5] = False
    assert bool(torch.isfinite(dl.float().permute(0, 2, 3, 1)[ok]).all())
    # the same labels under ignore_index=255 are fine
    loss2, dl2 = ops.cross_entropy_fused(logits, labels, 255)
    ref = torch.nn.functional.cross_entropy(logits.float(), labels, ignore_index=255)
    assert abs(float(loss2) - float(ref)) < 1e-2 and bool(torch.isfinite(dl2.float()).all())


def test_cross_entropy_fused_all_rows_ignored_is_nan_like_torch():
    from maskunet_b200 import ops
    logits = _cl(torch.randn(1, 19, 4, 8, device=DEV))
    labels = torch.full((1, 4, 8), 255, device=DEV)
    loss, dl = ops.cross_entropy_fused(logits, labels, 255)
    ref = torch.nn.functional.cross_entropy(logits, labels, ignore_index=255)
    assert bool(torch.isnan(ref)) and bool(torch.isnan(loss).all()) and float(dl.abs().max()) == 0.0


def test_unet_eval_mode_with_autograd_on_class_padded_head():
    """eval() with autograd enabled (frozen-BatchNorm fine-tuning, gradient checks): the class-padded head output takes
    the plain-torch affine BatchNorm path.  Logits equal the no_grad (fused inference kernels) forward within bf16
    tolerance and backward reaches every live parameter."""
    import maskunet_b200
    torch.manual_seed(0)
    net = maskunet_b200.UNet(3, 19, compute_dtype=torch.bfloat16, channels_last=True).to(DEV)
    net = net.to(memory_format=torch.channels_last).eval()
    x = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(1)).to(DEV)
    with torch.no_grad():
        ref = net(x).float()
    out = net(x)
    assert out.requires_grad and out.shape == (2, 19, 128, 128)
    assert float((out.float() - ref).norm() / ref.norm()) < 2e-2
    out.float().square().mean().backward()
    grads = [p.grad for n, p in net.named_parameters() if "emb_layer" not in n]
    assert all(g is not None and bool(torch.isfinite(g).all()) for g in grads)


def test_trainer_cuda_graph_step_matches_eager_steps():
    """Trainer(cuda_graph=True): after 3 eager steps the whole step (zero_grad, forward, fused CE, backward, AdamW) is
    captured once and replayed.  Same seeds, same batches: the loss sequence of the graph run follows the eager run
    (dropout disabled: eager and captured Philox offsets need not coincide), and the parameters keep moving."""
    import maskunet_b200
    from maskunet_b200.train import Trainer
    g = torch.Generator().manual_seed(3)
    xs = [torch.rand(2, 3, 128, 128, generator=g).to(DEV) for _ in range(6)]
    ys = [torch.randint(0, 19, (2, 128, 128), generator=g).to(DEV) for _ in range(6)]

    def run(graph):
        torch.manual_seed(11)
        net = maskunet_b200.UNet(3, 19, compute_dtype=torch.bfloat16, channels_last=True).to(DEV)
        net = net.to(memory_format=torch.channels_last).train()
        net.dropout.p = 0.0
        tr = Trainer(net, lr=1e-3, weight_decay=1e-2, cuda_graph=graph)
        losses, snaps = [], []
        for x, y in zip(xs, ys):
            losses.append(float(tr.step(x, y)))
            snaps.append(net.bottom2.conv_block[0].weight.detach().float().clone())
        return losses, tr, snaps

    eager, _, snaps_e = run(False)
    graph, tr, snaps_g = run(True)
    assert tr._graph is not None and tr.graph_launches > 300          # steps 4-6 were replays of the captured step
    print("eager", eager, "graph", graph)
    for a, b in zip(eager, graph):
        assert abs(a - b) < 5e-3 * abs(a), (eager, graph)
    assert eager[0] != eager[-1]
    # every replay is a real optimiser step: the parameters keep moving, by about as much as in the eager run (AdamW
    # normalises the update, so the noisy bf16 gradients make the two trajectories drift apart -- only sizes compare)
    for k in (3, 4, 5):
        d_g = float((snaps_g[k] - snaps_g[k - 1]).norm())
        d_e = float((snaps_e[k] - snaps_e[k - 1]).norm())
        assert d_g > 0 and 0.5 < d_g / d_e < 2.0, (k, d_g, d_e)
    with pytest.raises(RuntimeError):
        tr.step(xs[0][:1], ys[0][:1])                                  # the captured shapes are fixed
