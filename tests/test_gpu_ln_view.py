"""GPU checks of the fused re-view: the reference returns its [B, N, C] LayerNorm result RE-VIEWED as [B, C, H, W]
(ade_semantic.py:187-190, a .view, not a .permute).  A channels-last network holds that tensor as [B, H*W, C] memory --
a transpose of every sample's [C, N] view.  mu_residual_ln_fwd / _bwd with MU_X_TOKEN_MAJOR_VIEW write / read that layout
directly.  Bar: bit-identical to the token-major kernels followed (preceded) by the separate transpose pass."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)

SHAPES = [(2, 16384, 64), (3, 4096, 64), (2, 4096, 128), (3, 1024, 256), (2, 256, 256), (1, 1024, 128), (2, 64, 64)]


def _rand(*shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV).to(torch.bfloat16)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "b%d_n%d_c%d" % s)
def test_layernorm_view_forward_and_backward_bit_identical(shape):
    from maskunet_b200 import ops
    B, N, C = shape
    o, x, dy_tok = _rand(B, N, C, seed=1), _rand(B, N, C, seed=2), _rand(B, N, C, seed=3)
    gamma = torch.randn(C, device=DEV) * 0.5 + 1.0
    beta = torch.randn(C, device=DEV) * 0.1
    y, mean, rstd = ops.residual_ln_fwd(o, x, gamma, beta, 1e-5, True)
    want = ops.transpose(y.view(B, C, N))                       # [B, N, C]: channels-last memory of y.view(B, C, H, W)
    got, mean_v, rstd_v = ops.residual_ln_fwd(o, x, gamma, beta, 1e-5, True, True)
    assert torch.equal(got, want)
    assert torch.equal(mean_v, mean) and torch.equal(rstd_v, rstd)
    # the reference's own statement of the re-view
    ref = torch.nn.functional.layer_norm(o.float() + x.float(), (C,), gamma, beta, 1e-5)
    ref_view = ref.view(B, C, N).permute(0, 2, 1)               # channels-last memory of the [B, C, H, W] view
    assert (got.float() - ref_view).abs().max() < 6e-2

    # backward: the gradient of the view, channels-last -> token-major dy is its transpose back
    dy_view = ops.transpose(dy_tok.view(B, C, N))               # what the network hands back for dy_tok
    dz, delta, dg, db = ops.residual_ln_bwd(dy_tok, o, x, mean, rstd, gamma, True)
    dz_v, delta_v, dg_v, db_v = ops.residual_ln_bwd(dy_view, o, x, mean, rstd, gamma, True, True)
    assert torch.equal(dz_v, dz) and torch.equal(delta_v, delta)
    assert torch.allclose(dg_v, dg, rtol=1e-4, atol=1e-3) and torch.allclose(db_v, db, rtol=1e-4, atol=1e-3)


def test_layernorm_view_rejects_geometry_it_cannot_transpose():
    from maskunet_b200 import ops
    o, x = _rand(1, 96, 64, seed=1), _rand(1, 96, 64, seed=2)   # N % C != 0
    g = torch.ones(64, device=DEV)
    with pytest.raises(RuntimeError):
        ops.residual_ln_fwd(o, x, g, g, 1e-5, True, True)


@pytest.mark.parametrize("c_hw", [(64, 128), (128, 64), (256, 32), (256, 16)], ids=lambda s: "c%d_hw%d" % s)
def test_module_output_and_gradients_identical_with_and_without_fused_view(c_hw, monkeypatch):
    """Mask2FormerAttention on a channels-last bf16 activation: fused view on / off."""
    import maskunet_b200
    C, HW = c_hw
    torch.manual_seed(3)
    mod = maskunet_b200.Mask2FormerAttention(C, C).to(DEV)
    mod.compute_dtype = torch.bfloat16
    x0 = _rand(2, C, HW, HW, seed=4).contiguous(memory_format=torch.channels_last)
    gy = _rand(2, C, HW, HW, seed=5).contiguous(memory_format=torch.channels_last)
    maskunet_b200.set_deterministic(True)
    try:
        outs = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("MASKUNET_LN_VIEW", flag)
            mod.zero_grad(set_to_none=True)
            x = x0.clone(memory_format=torch.preserve_format).requires_grad_(True)
            y = mod(x)
            assert y.shape == x.shape and y.is_contiguous(memory_format=torch.channels_last)
            y.backward(gy)
            outs[flag] = (y.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in mod.named_parameters()})
    finally:
        maskunet_b200.set_deterministic(False)
    a, b = outs["1"], outs["0"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for n in a[2]:
        if n.startswith("norm."):      # dgamma / dbeta: same addends, another CTA partition
            assert torch.allclose(a[2][n], b[2][n], rtol=1e-4, atol=1e-3), n
        else:
            assert torch.equal(a[2][n], b[2][n]), n
