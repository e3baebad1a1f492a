"""CPU, build container only: the oracle against the reference's classes loaded live from /root/reference."""
import pytest
import torch

from conftest import rel_err
from oracle import mask_attention_oracle as mao
from oracle import unet_oracle as uo
from oracle.ref_loader import load_reference_classes, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 8, 8), (1, 128, 12, 20), (2, 256, 8, 8)])
def test_attention_module_live(B, C, H, W):
    ref = load_reference_classes("ade_semantic")
    torch.manual_seed(5)
    m = ref.Mask2FormerAttention(C, C)
    x = torch.randn(B, C, H, W, requires_grad=True)
    torch.manual_seed(11)
    y = m(x)
    torch.manual_seed(11)
    keep = mao.binarize_mask(mao.draw_mask_bits(B, H, W))          # same RNG stream -> identical bits
    assert torch.equal(keep, mao.keep_from_module_mask(m.mask))
    dy = torch.randn_like(y)
    (y * dy).sum().backward()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    out = mao.attention_forward(x.detach(), sd, keep)
    assert rel_err(mao.module_output(out["y"], C, H, W), y) < 2e-6
    g = mao.attention_backward(out, dy.view(B, H * W, C))
    assert rel_err(g["x"].view_as(x), x.grad) < 5e-6
    for k, p in m.named_parameters():
        if k != "key.bias":
            assert rel_err(g[k], p.grad) < 5e-6, k


def test_mask_cache_semantics_of_reference():
    """Appendix A.3: cached after first forward, no RNG afterwards, different batch later errors."""
    ref = load_reference_classes("ade_semantic")
    m = ref.Mask2FormerAttention(64, 64)
    m(torch.randn(2, 64, 4, 4))
    first = m.mask
    state = torch.get_rng_state()
    x2 = torch.randn(2, 64, 4, 4)
    after_input = torch.get_rng_state()
    m(x2)
    assert m.mask is first                                   # cached, not redrawn
    assert torch.equal(torch.get_rng_state(), after_input)   # second forward consumes no RNG
    assert not torch.equal(state, after_input)
    with pytest.raises(RuntimeError):
        m(torch.randn(3, 64, 4, 4))


@pytest.mark.parametrize("script,variant,c_out", [("ade_semantic", "semantic", 150), ("coco_panoptic", "semantic", 133),
                                                  ("city_instance", "instance", 19)])
def test_unet_init_and_keys_identical(script, variant, c_out):
    ref = load_reference_classes(script)
    torch.manual_seed(3)
    model = ref.UNet(3, c_out)
    torch.manual_seed(3)
    sd = uo.init_state(3, c_out, variant)
    rsd = model.state_dict()
    assert list(rsd.keys()) == list(sd.keys())
    assert all(torch.equal(rsd[k], sd[k]) for k in sd)


def test_mean_iou_live():
    from oracle import metrics_oracle as mo
    from oracle.ref_loader import load_reference_functions
    fn = load_reference_functions("ade_semantic", ("mean_iou",)).mean_iou
    g = torch.Generator().manual_seed(3)
    for C in (5, 150):
        logits = torch.relu(torch.randn(2, C, 12, 12, generator=g))
        labels = torch.randint(0, C, (2, 12, 12), generator=g)
        assert abs(float(fn(logits, labels, C)) - float(mo.mean_iou(logits, labels, C))) < 1e-6
        assert torch.equal(torch.argmax(torch.softmax(logits / 0.5, dim=1), dim=1), mo.class_map(logits))


@pytest.mark.parametrize("script,ignore", [("coco_panoptic", None), ("city_instance", 255), ("ade_panoptic", None),
                                           ("city_panoptic", None)])
def test_instance_contrastive_loss_live(script, ignore):
    """The restated InstanceContrastiveLoss against every copy of the reference's class, same CPU generator stream."""
    from oracle import instance_loss_oracle as ilo
    paths = {"ade_panoptic": "code/ade20k/ade_panoptic.py", "city_panoptic": "code/cityscapes/city_panoptic.py"}
    Ref = load_reference_classes(paths.get(script, script), ("InstanceContrastiveLoss",)).InstanceContrastiveLoss
    g = torch.Generator().manual_seed(5)
    for B, C, H, W in ((2, 5, 8, 8), (4, 19, 16, 16), (3, 7, 8, 12)):
        sem = torch.randn(B, C, H, W, generator=g).requires_grad_()
        im = torch.randint(0, 6, (B, H, W), generator=g) * 37
        im[0, 0, :3] = 255
        im[1, 2, 3] = 99999
        torch.manual_seed(77)
        ref = Ref()(sem, im)
        ref.backward()
        gref, sem.grad = sem.grad.clone(), None
        state = torch.get_rng_state()
        torch.manual_seed(77)
        ours, _ = ilo.instance_contrastive_loss(sem, im, 1.0, ignore)
        ours.backward()
        assert torch.equal(state, torch.get_rng_state())
        assert abs(float(ref) - float(ours)) < 1e-6 and rel_err(sem.grad, gref) < 1e-5
    # nothing qualifies: background only
    assert float(ilo.instance_contrastive_loss(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))[0]) == 0.0
