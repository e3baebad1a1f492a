"""CPU: the C-ABI library loads and exports what include/maskunet_b200.h declares; drop-in surface checks."""
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    from maskunet_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "maskunet_b200.h")).read()
    declared = set(re.findall(r"\b(mu_[a-z0-9_]+)\s*\(", header))
    assert {"mu_version", "mu_last_error", "mu_mask_binarize", "mu_attn_fwd", "mu_attn_bwd"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    for name in _lib.SIGNATURES:
        assert name in declared, f"{name} bound in _lib.py but not declared in the header"
    assert lib.mu_version() >= 100


def test_ops_are_registered_and_refuse_cpu_tensors():
    import maskunet_b200  # noqa: F401
    for name in ("mask_binarize", "attn_fwd", "attn_bwd", "mask_attention", "mask_attention_bwd", "qkv_project",
                 "residual_ln_fwd", "residual_ln_bwd", "qkv_project_bwd", "transpose"):
        assert hasattr(torch.ops.maskunet, name)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.maskunet.mask_binarize(torch.zeros(1, 8, dtype=torch.int64))


def test_fake_tensor_shapes():
    import maskunet_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(2, 400, 64, device="cuda", dtype=torch.bfloat16)   # token-major
        w = torch.empty(192, 64, device="cuda")
        b = torch.empty(192, device="cuda")
        g = torch.empty(64, device="cuda")
        rank = torch.empty(2, 400, dtype=torch.int32, device="cuda")
        nk = torch.empty(2, dtype=torch.int32, device="cuda")
        outs = torch.ops.maskunet.mask_attention(x, w, b, g, g, rank, rank, nk, 1e-5, True)
        assert outs[0].shape == (2, 400, 64) and outs[2].shape == (2, 512, 64) and outs[5].dtype == torch.float32


def test_module_surface_matches_reference_contract():
    from maskunet_b200 import InstanceUNet, Mask2FormerAttention, UNet
    m = Mask2FormerAttention(64, 64)
    assert list(m.state_dict().keys()) == ["query.weight", "query.bias", "key.weight", "key.bias", "value.weight",
                                           "value.bias", "norm.weight", "norm.bias"]
    assert m.mask is None and m.channels == 64 and m.size == 64
    with pytest.raises(ValueError, match="Input channel size does not match initialized channel size."):
        m(torch.zeros(1, 32, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 64, 4, 4))
    for cls, fname, c_out in ((UNet, "unet_semantic", 150), (InstanceUNet, "unet_instance", 19)):
        with open(os.path.join(GOLDEN, fname + ".json")) as fh:
            meta = json.load(fh)
        torch.manual_seed(meta["seed"])
        net = cls(3, c_out)
        names = [k for k, _ in net.named_parameters()]
        assert names == meta["grad_names"]                      # same parameters, same order as the reference
        digest = {k: float(v.double().abs().sum()) for k, v in net.state_dict().items() if v.dtype.is_floating_point}
        assert set(digest) == set(meta["param_digest"])           # same state_dict keys (incl. dead emb_layer.*)
        same_rng = all(abs(digest[k] - v) <= 1e-6 * max(1.0, abs(v)) for k, v in meta["param_digest"].items())
        assert same_rng, "initialisation does not reproduce the reference's RNG order"


def test_widened_ops_are_registered_with_fake_shapes():
    """The ops of the rows widened after the hot path exist, infer shapes without a device and have no
    CPU kernel."""
    import maskunet_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("instance_triplet", "instance_triplet_bwd", "argmax_iou", "to_tensor_u8", "query_mask_bits",
                 "query_attn_fwd", "query_attn_bwd"):
        assert hasattr(torch.ops.maskunet, name), name
    with FakeTensorMode():
        bf = dict(device="cuda", dtype=torch.bfloat16)
        qe, feat = torch.empty(2, 100, 256, **bf), torch.empty(2, 1000, 256, **bf)
        bits, bits_t, cnt, logits = torch.ops.maskunet.query_mask_bits(qe, feat, False)
        assert bits.shape == (2, 100, 32) and bits_t.shape == (2, 1024, 4) and cnt.shape == (2, 100) and logits.numel() == 0
        qh, kh = torch.empty(8, 100, 64, **bf), torch.empty(8, 1024, 64, **bf)
        o, lse = torch.ops.maskunet.query_attn_fwd(qh, kh, kh, bits, 4, 1000, 0.125)
        assert o.shape == qh.shape and lse.shape == (8, 100) and lse.dtype == torch.float32
        u8 = torch.empty(4, 128, 128, 3, device="cuda", dtype=torch.uint8)
        x = torch.ops.maskunet.to_tensor_u8(u8, True, True, 8)
        assert x.shape == (4, 8, 128, 128) and x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last)
        sem = torch.empty(2, 19, 16, 16, device="cuda")
        loss, sel, dist = torch.ops.maskunet.instance_triplet(sem, torch.empty(512, device="cuda", dtype=torch.int64),
                                                              torch.empty(3, 5, device="cuda", dtype=torch.int64), 1.0)
        assert loss.shape == (1,) and sel.shape == (5, 6) and dist.shape == (5, 2)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.maskunet.to_tensor_u8(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), False, False, 0)
    crit = maskunet_b200.InstanceContrastiveLoss(margin=1.0, ignore_value=255)
    assert crit.margin == 1.0 and crit.ignore_value == 255 and len(list(crit.parameters())) == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))


def test_c_abi_error_convention_without_a_device():
    """SURVEY 8(b) error convention: 0 on success, a negative mu_status for argument errors (checked before anything
    touches the device), text behind mu_last_error(); a tcgen05 entry point on a machine without an sm_100 device
    answers MU_ERR_ARCH instead of crashing."""
    import ctypes
    from maskunet_b200 import _lib
    lib = _lib.load()
    MU_ERR_BAD_SHAPE, MU_ERR_BAD_DTYPE, MU_ERR_MISALIGNED, MU_ERR_ARCH, MU_ERR_NULL = -1, -2, -3, -4, -5
    buf = ctypes.create_string_buffer(4096)
    base = (ctypes.addressof(buf) + 15) & ~15            # a 16-byte aligned (host) address: never dereferenced here
    P = ctypes.c_void_p
    err = lambda: lib.mu_last_error().decode()
    # channels must be 64 / 128 / 256
    rc = lib.mu_attn_fwd(P(base), P(base), P(base), P(base), P(base), P(base), 1, 128, 128, 65, _lib.MU_BF16, P(0))
    assert rc == MU_ERR_BAD_SHAPE and "channels must be 64, 128 or 256" in err()
    # NKP must be a multiple of 128 and >= N
    rc = lib.mu_attn_fwd(P(base), P(base), P(base), P(base), P(base), P(base), 1, 200, 128, 64, _lib.MU_BF16, P(0))
    assert rc == MU_ERR_BAD_SHAPE and "NKP" in err()
    rc = lib.mu_attn_fwd(P(base), P(base), P(base), P(base), P(base), P(base), 1, 128, 128, 64, 7, P(0))
    assert rc == MU_ERR_BAD_DTYPE and "dtype" in err()
    rc = lib.mu_attn_fwd(P(0), P(base), P(base), P(base), P(base), P(base), 1, 128, 128, 64, _lib.MU_BF16, P(0))
    assert rc == MU_ERR_NULL and "null pointer" in err()
    rc = lib.mu_attn_fwd(P(base + 4), P(base), P(base), P(base), P(base), P(base), 1, 128, 128, 64, _lib.MU_BF16, P(0))
    assert rc == MU_ERR_MISALIGNED and "16-byte" in err()
    rc = lib.mu_attn_bwd(*([P(base)] * 11), P(base), 0, 1, 128, 128, 64, _lib.MU_BF16, P(0))
    assert rc != 0                                        # workspace too small (or no device): never a crash
    # the order semaphores of deterministic mode (int32 per sample, channel half, 64-query tile) and, for 128 / 256
    # channels, the fp32 dQ accumulator (64 channels add bf16 partial tiles straight into dq)
    # (64 channels: instead of the accumulator, the 32-byte-per-query statistics slab of the folded lse / delta, rows padded to 128)
    assert lib.mu_attn_bwd_workspace_bytes(2, 400, 64, _lib.MU_BF16) == 2 * 512 * 32 + 2 * 2 * 7 * 4 + 16
    assert lib.mu_attn_bwd_workspace_bytes(2, 400, 128, _lib.MU_BF16) == 2 * 400 * 128 * 4 + 2 * 2 * 7 * 4 + 16
    assert lib.mu_get_deterministic() == 0
    lib.mu_set_deterministic(1)
    assert lib.mu_get_deterministic() == 1
    lib.mu_set_deterministic(0)
    assert lib.mu_query_attn_bwd_workspace_bytes(8, 100, 64) >= 16      # 64-wide heads: bf16 partial tiles go into dq
    rc = lib.mu_query_mask_bits(P(base), P(base), 1, 100, 1000, 256, P(base), P(base), P(base), P(0), _lib.MU_F32, P(0))
    assert rc == MU_ERR_BAD_DTYPE
    rc = lib.mu_instance_triplet_fwd(P(base), P(0), 1, 2, 4, 4, P(base), P(base), 1, 1.0, 1e-6, P(base), P(base), P(base),
                                     _lib.MU_F32, P(0))
    assert rc == MU_ERR_NULL
    if not torch.cuda.is_available():                     # build container: the tcgen05 path refuses to run
        assert lib.mu_device_supported() == 0
        rc = lib.mu_attn_fwd(P(base), P(base), P(base), P(base), P(base), P(base), 1, 128, 128, 64, _lib.MU_BF16, P(0))
        assert rc == MU_ERR_ARCH and "sm_100" in err()


def test_eval_mode_affine_batchnorm_on_a_class_padded_tensor():
    """The plain-torch path used for the class-padded head output in eval() with autograd on: equals
    nn.BatchNorm2d.eval() on the real channels, keeps the pad channels zero, and is differentiable."""
    from maskunet_b200.modules import _bn_eval_affine
    g = torch.Generator().manual_seed(0)
    bn = torch.nn.BatchNorm2d(19).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(19, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(19, generator=g))
        bn.running_mean.copy_(torch.randn(19, generator=g))
        bn.running_var.copy_(torch.rand(19, generator=g) + 0.1)
    x = torch.zeros(2, 32, 5, 7)
    x[:, :19] = torch.randn(2, 19, 5, 7, generator=g)
    x = x.contiguous(memory_format=torch.channels_last).requires_grad_()
    y = _bn_eval_affine(x, bn, 13)
    assert torch.allclose(y[:, :19], bn(x[:, :19]), atol=1e-6) and float(y[:, 19:].abs().max()) == 0.0
    y.square().sum().backward()
    assert x.grad is not None and bn.weight.grad is not None and float(x.grad[:, 19:].abs().max()) == 0.0
