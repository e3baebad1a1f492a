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
