"""CPU: the oracle restatement against golden vectors produced by the reference's own classes."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ATTN_CASES, GOLDEN, load_attn_golden, rel_err
from oracle import mask_attention_oracle as mao
from oracle import unet_oracle as uo


@pytest.mark.parametrize("name", ATTN_CASES)
def test_attention_forward_backward_matches_reference_golden(name):
    g = load_attn_golden(name)
    x = g["x"]
    B, C, H, W = x.shape
    out = mao.attention_forward(x, g["params"], g["keep"])
    y = mao.module_output(out["y"], C, H, W)
    assert rel_err(y, g["y"]) < 2e-6
    grads = mao.attention_backward(out, g["dy"].view(B, H * W, C))
    assert rel_err(grads["x"].view_as(x), g["dx"]) < 5e-6
    scale = max(float(v.abs().max()) for v in g["grads"].values())
    for k, ref in g["grads"].items():
        if k == "key.bias":  # analytically zero (softmax is shift invariant): absolute check
            assert float((grads[k] - ref).abs().max()) < 1e-5 * scale
        else:
            assert rel_err(grads[k], ref) < 5e-6, k


def test_kat_scalars_from_survey():
    # B2 C64 8x8 -> keep [25, 30], y.abs().sum() = 6566.9102
    g = load_attn_golden("attn_b2_c64_8x8")
    assert g["keep"].sum(1).tolist() == [25, 30]
    assert abs(float(g["kat"][0]) - 6566.9102) < 1e-2
    g = load_attn_golden("attn_b1_c256_16x16")
    assert g["keep"].sum(1).tolist() == [123]
    assert abs(float(g["kat"][0]) - 52385.5781) < 1e-1


def test_binarize_is_bits_gt_half():
    bits = torch.tensor([[0, 1, 1, 0, 2, -1]])
    keep = mao.binarize_mask(bits)
    assert keep.tolist() == [[False, True, True, False, True, False]]
    bias = mao.additive_bias(keep)
    assert bias[0, 0] == float("-inf") and bias[0, 1] == 0
    m = mao.expand_bias(bias, 6)
    assert m.shape == (1, 6, 6) and m.stride() == (6, 0, 1)


def test_mask_draw_consumes_rng_like_reference():
    torch.manual_seed(7)
    a = mao.draw_mask_bits(2, 4, 4)
    torch.manual_seed(7)
    b = torch.randint(0, 2, (2, 4, 4))
    assert torch.equal(a, b)


def _unet_golden(fname):
    with open(os.path.join(GOLDEN, fname + ".json")) as fh:
        meta = json.load(fh)
    z = np.load(os.path.join(GOLDEN, fname + ".npz"))
    return meta, z


@pytest.mark.parametrize("fname,variant", [("unet_semantic", "semantic"), ("unet_instance", "instance")])
def test_unet_oracle_matches_reference_golden(fname, variant):
    meta, z = _unet_golden(fname)
    torch.manual_seed(meta["seed"])
    sd = uo.init_state(3, meta["c_out"], variant)
    digest = {k: float(v.double().abs().sum()) for k, v in sd.items() if v.dtype.is_floating_point}
    bad = [k for k, v in meta["param_digest"].items() if abs(digest[k] - v) > 1e-6 * max(1.0, abs(v))]
    # the driver's boxes run the image the golden was made in: a different RNG stream is an error, not a skip
    assert not bad, f"initial parameters differ from the golden's (torch RNG stream changed?): {bad[:3]}"
    x = torch.rand(meta["batch"], 3, 128, 128)
    keeps = {}
    for n, _, side in uo.ATTN_SITES:
        packed = torch.from_numpy(z[f"keep.{n}"])
        keeps[n] = torch.from_numpy(np.unpackbits(packed.numpy(), axis=1)[:, : side * side]).bool()
    with torch.no_grad():
        outs = uo.unet_forward(sd, x, keeps, variant=variant)
    outs = outs if isinstance(outs, tuple) else (outs,)
    for i, o in enumerate(outs):
        idx = torch.from_numpy(z[f"out{i}.sample_idx"])
        ref = torch.from_numpy(z[f"out{i}.sample"])
        got = o.reshape(-1)[idx]
        assert float((got - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max())), i
    am = outs[0].argmax(1).numpy().astype(np.uint8)
    margin = z["argmax_margin"].astype(np.float32)
    safe = margin > 1e-3
    assert (am[safe] == z["argmax"][safe]).all()
    assert (am == z["argmax"]).mean() > 0.999


@pytest.mark.parametrize("name,C", [("postproc_c150", 150), ("postproc_c19", 19)])
def test_metrics_oracle_matches_reference_golden(name, C):
    """oracle/metrics_oracle.py against outputs of the reference's own mean_iou (ade_semantic.py:128-146)."""
    import os
    import numpy as np
    import torch
    from conftest import GOLDEN
    from oracle import metrics_oracle as mo
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    logits, labels = torch.from_numpy(z["logits"]), torch.from_numpy(z["labels"])
    assert torch.equal(mo.class_map(logits), torch.from_numpy(z["pred"].astype(np.int64)))   # bit-exact class map
    assert abs(float(mo.mean_iou(logits, labels, C)) - float(z["miou"][0])) < 1e-6


INSTANCE_LOSS_CASES = ["instance_loss_coco", "instance_loss_city", "instance_loss_blobs"]


@pytest.mark.parametrize("name", INSTANCE_LOSS_CASES)
def test_instance_contrastive_loss_matches_reference_golden(name):
    """oracle/instance_loss_oracle.py against the reference's own InstanceContrastiveLoss (coco_panoptic.py:482-521,
    city_instance.py:279-307): loss, gradient, and the same number of draws from the CPU generator."""
    from oracle import instance_loss_oracle as ilo
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    seed, ignore, next_draw = (int(v) for v in z["meta"])
    sem = torch.from_numpy(z["sem"]).requires_grad_()
    torch.manual_seed(seed)
    loss, sel = ilo.instance_contrastive_loss(sem, torch.from_numpy(z["instance_mask"]), 1.0,
                                              None if ignore < 0 else ignore)
    assert int(torch.randint(0, 2 ** 31, (1,))) == next_draw       # generator left in the reference's state
    assert abs(float(loss) - float(z["loss"][0])) < 1e-6
    loss.backward()
    assert rel_err(sem.grad, torch.from_numpy(z["grad"])) < 1e-5
    assert len(sel) >= 3


def test_to_tensor_matches_torchvision_golden():
    """oracle/input_oracle.py against torchvision's own ToTensor (the reference's transform, ade_semantic.py:9,85):
    all 256 byte values and a random RGB image, bit for bit."""
    from oracle import input_oracle as io_
    z = np.load(os.path.join(GOLDEN, "to_tensor.npz"))
    for k in ("ramp", "img"):
        got = io_.to_tensor(z[k][None])[0]
        assert got.dtype == np.float32 and np.array_equal(got, z[k + "_out"])


def test_generalised_mode_oracle_agrees_with_torch_sdpa():
    """oracle/query_attention_oracle.py is builder-written (parity unpinned by reference: the reference has no such
    stage).  It is at least held to PyTorch's own masked attention and to the sigmoid rule it states."""
    import torch.nn.functional as F
    from oracle import query_attention_oracle as qo
    g = torch.Generator().manual_seed(2)
    B, Q, N, C, heads = 2, 37, 200, 64, 4
    qe, feat = torch.randn(B, Q, C, generator=g), torch.randn(B, N, C, generator=g)
    qe[0, 3] = 0                                                     # all-zero logits: keeps nothing -> attends everything
    logits = qo.mask_logits(qe, feat)
    keep, count = qo.keep_from_logits(logits)
    assert int(count[0, 3]) == 0 and bool(keep[0, 3].all())
    raw = torch.sigmoid(logits) > 0.5
    assert torch.equal(keep[count > 0], raw[count > 0])
    # the fp32 boundary the kernel implements: sigmoid(x) > 0.5  <=>  x > 1.5 * 2^-24
    t = torch.tensor([0.0, 1.5 * 2.0 ** -24, float(np.nextafter(np.float32(1.5 * 2.0 ** -24), np.float32(1))), 2.0 ** -23])
    assert (torch.sigmoid(t) > 0.5).tolist() == [False, False, True, True]
    assert torch.equal(torch.sigmoid(t) > 0.5, t > 1.5 * 2.0 ** -24)
    q, k, v = (torch.randn(B, n, C, generator=g) for n in (Q, N, N))
    out = qo.attention(q, k, v, keep, heads)
    d = C // heads
    split = lambda x: x.view(B, -1, heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(split(q), split(k), split(v), attn_mask=keep.unsqueeze(1))
    assert rel_err(out, ref.transpose(1, 2).reshape(B, Q, C)) < 1e-5
    words = (N + 127) // 128 * 4
    assert torch.equal(qo.unpack_bits(qo.pack_bits(keep, words), N), keep)


@pytest.mark.parametrize("name", ["ade_like", "exact_2x", "enlarge", "wide", "tall_to_rect"])
def test_resize_oracle_matches_cv2_golden(name):
    """oracle/resize_oracle.py against cv2's own INTER_LINEAR / INTER_NEAREST outputs (ade_semantic.py:72-73)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    from oracle import resize_oracle as ro
    z = np.load(os.path.join(GOLDEN, "resize.npz"))
    size = z[name + ".linear"].shape[:2]
    assert np.array_equal(ro.resize_linear_u8(z[name + ".img"], size), z[name + ".linear"])
    assert np.array_equal(ro.resize_nearest_u8(z[name + ".mask"], size), z[name + ".nearest"])
