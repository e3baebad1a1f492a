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


def test_cpu_baseline_step_is_the_reference_step_value_and_speed():
    """bench.py's CPU arm (OracleTrainer(reference_ops=True)) against the reference's own classes + CrossEntropyLoss +
    AdamW(lr 5e-5, wd 1e-1) (ade_semantic.py:377-379, :394-401) on the same cores: the same losses step by step
    (identical initial state, masks and dropout stream) and the same speed (within 15 %; measured 0.95-1.0x -- the
    explicit parity restatement, which materialises max / exp / sum / div N x N tensors, costs ~1.7-1.9x and must never
    be what the baseline times)."""
    import time
    ref = load_reference_classes("ade_semantic")
    B = 1
    img = torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(0))
    lab = torch.randint(0, 150, (B, 128, 128), generator=torch.Generator().manual_seed(1))

    torch.manual_seed(42)
    model = ref.UNet(3, 150)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=1e-1)
    crit = torch.nn.CrossEntropyLoss()

    def ref_step():
        opt.zero_grad()
        loss = crit(model(img), lab)
        loss.backward()
        opt.step()
        return float(loss)

    port = uo.OracleTrainer(3, 150, lr=5e-5, weight_decay=1e-1, seed=42, dropout_p=0.3, reference_ops=True)
    # same RNG stream for masks + dropout in both arms: reseed before each arm's first step
    torch.manual_seed(7)
    l_ref = [ref_step()]
    torch.manual_seed(7)
    l_port = [port.step(img, lab)]
    assert abs(l_ref[0] - l_port[0]) < 1e-4 * abs(l_ref[0]), (l_ref, l_port)
    t_ref, t_port = [], []
    for _ in range(2):                       # interleaved, best of two: robust against a noisy neighbour
        t = time.perf_counter(); ref_step(); t_ref.append(time.perf_counter() - t)
        t = time.perf_counter(); port.step(img, lab); t_port.append(time.perf_counter() - t)
    ratio = min(t_port) / min(t_ref)
    print(f"CPU baseline port / reference step time: {ratio:.3f} ({min(t_port):.2f} s vs {min(t_ref):.2f} s, batch {B})")
    assert 0.8 < ratio < 1.15, (t_port, t_ref)


def test_resize_oracle_fuzz_against_cv2():
    """oracle/resize_oracle.py against the third-party library the reference calls (cv2.resize, ade_semantic.py:72-73),
    live: random source sizes, down- and up-scaling, 1 and 3 channels, three output sizes -- every byte equal."""
    cv2 = pytest.importorskip("cv2")
    import numpy as np
    from oracle import resize_oracle as ro
    rng = np.random.default_rng(0)
    shapes = [(512, 683), (1024, 2048), (480, 640), (256, 256), (128, 128), (100, 37), (1, 5), (2000, 3)]
    shapes += [tuple(int(v) for v in rng.integers(1, 700, 2)) for _ in range(25)]
    for sh in shapes:
        for oh, ow in ((128, 128), (64, 96), (200, 150)):
            img = rng.integers(0, 256, size=(sh[0], sh[1], 3), dtype=np.uint8)
            assert np.array_equal(cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR),
                                  ro.resize_linear_u8(img, (oh, ow))), (sh, oh, ow)
            m = img[:, :, 0]
            assert np.array_equal(cv2.resize(m, (ow, oh), interpolation=cv2.INTER_NEAREST),
                                  ro.resize_nearest_u8(m, (oh, ow))), (sh, oh, ow)
            assert np.array_equal(cv2.resize(m, (ow, oh), interpolation=cv2.INTER_LINEAR),
                                  ro.resize_linear_u8(m, (oh, ow))), (sh, oh, ow)
