"""GPU checks of the skip-gradient folding: where the reference's graph gives an activation two consumers -- the residual
ConvBlock gelu(x + block(x)) (ade_semantic.py:205-208) and the U-Net skip connections x1..x3 (:301-312) -- autograd adds
the two gradients in a separate pass.  Here the second consumer's gradient reaches the first consumer's backward kernel
(ops._Conv3x3Skip / ops._MaxPool2Skip) and the kernel adds into it: the TMA unit's bf16 reduce-add for the convolution
data gradient, a read-modify-write in the pooling backward.  Bar: BIT-IDENTICAL to gradient-then-add (one fp32 add of two
bf16 values, rounded once, every element touched exactly once)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)

CASES = [(2, 64, 64, 128, 128), (2, 128, 128, 128, 128), (2, 128, 64, 4, 128), (3, 128, 128, 64, 64),
         (2, 256, 256, 32, 32), (2, 512, 512, 16, 16), (4, 64, 64, 64, 64), (2, 64, 64, 8, 32)]


def _cl(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "b%d_%dto%d_%dx%d" % c)
@pytest.mark.parametrize("pair", ["0", "2"])
def test_conv_data_gradient_added_in_place_is_bit_identical(case, pair, monkeypatch):
    from maskunet_b200 import ops
    monkeypatch.setenv("MU_CONV_PAIR", pair)               # plain CTAs / CTA pairs forced
    B, Cin, Cout, H, W = case
    dy = _cl(B, Cout, H, W, seed=1)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) / (3.0 * Cin ** 0.5)
    _, wd = ops.conv_prep_weights(w, True)
    other = _cl(B, Cin, H, W, seed=2, scale=0.7)
    dx = ops.conv3x3_bwd_data(dy, wd)
    want = other + dx                                       # autograd's accumulation: fp32 add, one rounding
    got = other.clone(memory_format=torch.preserve_format)
    ops.conv3x3_bwd_data_acc(dy, wd, got)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", [(2, 64, 128, 128), (3, 128, 64, 64), (2, 256, 32, 32), (1, 8, 4, 6)])
def test_maxpool_gradient_added_in_place_is_bit_identical(shape):
    from maskunet_b200 import ops
    B, C, H, W = shape
    x = _cl(B, C, H, W, seed=3)
    dy = _cl(B, C, H // 2, W // 2, seed=4)
    other = _cl(B, C, H, W, seed=5)
    want = other + ops.maxpool2_bwd(x, dy)
    got = other.clone(memory_format=torch.preserve_format)
    ops.maxpool2_bwd_acc(x, dy, got)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


def test_residual_block_and_skip_functions_match_autograd_accumulation():
    """The autograd plumbing: gradients of x and the weights through (conv -> x_skip consumer) equal the unfolded graph."""
    from maskunet_b200 import ops
    B, C, H, W = 2, 64, 64, 64
    x0 = _cl(B, C, H, W, seed=6)
    w0 = torch.randn(C, C, 3, 3, device=DEV) / (3.0 * C ** 0.5)
    gy, gs = _cl(B, C, H, W, seed=7), _cl(B, C, H, W, seed=8)

    def run(fold):
        x = x0.clone(memory_format=torch.preserve_format).requires_grad_(True)
        w = w0.clone().requires_grad_(True)
        if fold:
            y, _, xs = ops.conv3x3_skip(x, w, False)
        else:
            y, xs = ops.conv3x3(x, w, False)[0], x
        z = xs * 1.5                                        # the second consumer
        torch.autograd.backward([y, z], [gy, gs])
        return x.grad, w.grad

    import maskunet_b200
    maskunet_b200.set_deterministic(True)                   # (the weight gradient's split-K sums in a fixed order)
    try:
        (dx_a, dw_a), (dx_b, dw_b) = run(True), run(False)
    finally:
        maskunet_b200.set_deterministic(False)
    assert torch.equal(dx_a, dx_b)
    assert torch.equal(dw_a, dw_b)

    def run_pool(fold):
        x = x0.clone(memory_format=torch.preserve_format).requires_grad_(True)
        if fold:
            y, xs = ops.maxpool2_skip(x)
        else:
            y, xs = ops.maxpool2(x), x
        z = xs * 0.5
        torch.autograd.backward([y, z], [gy[:, :, : H // 2, : W // 2].contiguous(memory_format=torch.channels_last), gs])
        return x.grad

    assert torch.equal(run_pool(True), run_pool(False))
    # only one of the two outputs used: the gradient still arrives
    x = x0.clone(memory_format=torch.preserve_format).requires_grad_(True)
    y, _, xs = ops.conv3x3_skip(x, w0.clone().requires_grad_(True), False)
    (xs * 2.0).sum().backward()
    assert torch.equal(x.grad, torch.full_like(x0, 2.0))


def test_unet_gradients_identical_with_and_without_folding(monkeypatch):
    """Whole bf16 channels-last U-Net train step, deterministic mode (bit-reproducible reductions): the folded graph's
    loss and parameter gradients are bit-identical to the graph in which autograd accumulates."""
    import maskunet_b200
    from maskunet_b200.train import Trainer
    torch.manual_seed(5)
    net = maskunet_b200.UNet(3, 150, compute_dtype=torch.bfloat16, channels_last=True).to(DEV)
    net = net.to(memory_format=torch.channels_last).train()
    g = torch.Generator().manual_seed(6)
    x = torch.rand(2, 3, 128, 128, generator=g).to(DEV)
    y = torch.randint(0, 150, (2, 128, 128), generator=g).to(DEV)
    tr = Trainer(net)

    def run(fold):
        monkeypatch.setenv("MASKUNET_FOLD_SKIP_GRADS", fold)
        torch.manual_seed(99)                       # dropout mask
        net.zero_grad(set_to_none=True)
        loss = tr.forward_backward(x, y)
        return loss.detach().clone(), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

    run("1")                                        # draws the attention masks
    maskunet_b200.set_deterministic(True)
    try:
        a, b = run("1"), run("0")
    finally:
        maskunet_b200.set_deterministic(False)
    assert torch.equal(a[0], b[0]) and len(a[1]) == len(b[1]) > 100
    differing = [n for n in a[1] if not torch.equal(a[1][n], b[1][n])]
    assert not differing, differing[:8]
