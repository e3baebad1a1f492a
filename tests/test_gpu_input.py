"""GPU parity of the device ToTensor against torchvision's own output (goldens) and the oracle:
bit-exact in fp32; bf16 = the fp32 value rounded once; the padded channels-last image feeds the U-Net unchanged."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import input_oracle as io_

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def test_to_tensor_bit_exact_against_torchvision_golden():
    from maskunet_b200 import data
    z = np.load(os.path.join(GOLDEN, "to_tensor.npz"))
    for k in ("ramp", "img"):
        u8 = torch.from_numpy(z[k])[None].to(DEV)
        want = torch.from_numpy(z[k + "_out"])[None]
        got = data.to_tensor(u8)
        assert got.dtype == torch.float32 and got.is_contiguous() and torch.equal(got.cpu(), want)
        cl = data.to_tensor(u8, torch.float32, channels_last=True)
        assert cl.is_contiguous(memory_format=torch.channels_last) and torch.equal(cl.cpu(), want)
        bf = data.to_tensor(u8, torch.bfloat16, channels_last=True, pad_to=8)
        assert bf.shape[1] == 8 and bf.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(bf[:, :3].float().cpu(), want.to(torch.bfloat16).float())
        assert float(bf[:, 3:].abs().max()) == 0.0


def test_full_size_batch_and_errors():
    from maskunet_b200 import data
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (64, 128, 128, 3), dtype=torch.uint8, generator=g)
    got = data.to_tensor(u8.to(DEV))
    assert torch.equal(got.cpu(), torch.from_numpy(io_.to_tensor(u8.numpy())))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        data.to_tensor(u8)
    with pytest.raises(ValueError):
        data.to_tensor(u8.to(DEV), pad_to=8)


def test_padded_image_feeds_the_unet_unchanged_and_prefetcher_order():
    import maskunet_b200
    from maskunet_b200 import data
    torch.manual_seed(0)
    net = maskunet_b200.UNet(3, 19, compute_dtype=torch.bfloat16, channels_last=True).to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randint(0, 256, (2, 128, 128, 3), dtype=torch.uint8, generator=g),
                torch.randint(0, 19, (2, 128, 128), generator=g)) for _ in range(3)]
    with torch.no_grad():
        for (u8, lab), (x, y) in zip(batches, data.DevicePrefetcher(batches, DEV)):
            assert x.shape == (2, 8, 128, 128) and x.dtype == torch.bfloat16 and torch.equal(y.cpu(), lab)
            ref_in = torch.from_numpy(io_.to_tensor(u8.numpy())).to(DEV)          # what the reference would feed
            assert torch.equal(net(x), net(ref_in))


# ------------------------------------------------------------------ cv2.resize on the device (ade_semantic.py:72-73)
RESIZE_CASES = ["ade_like", "exact_2x", "enlarge", "wide", "tall_to_rect"]


def test_resize_bit_exact_against_cv2_golden():
    """INTER_LINEAR image + ToTensor and INTER_NEAREST label map against cv2's own outputs (tests/golden/resize.npz),
    every source size in ONE batch call (sizes differ per image)."""
    from maskunet_b200 import data
    z = np.load(os.path.join(GOLDEN, "resize.npz"))
    groups = {}
    for name in RESIZE_CASES:
        groups.setdefault(tuple(z[name + ".linear"].shape[:2]), []).append(name)
    for size, names in groups.items():
        imgs = [torch.from_numpy(z[n + ".img"]).to(DEV) for n in names]
        masks = [torch.from_numpy(z[n + ".mask"]).to(DEV) for n in names]
        raw = data.resize_to_tensor(imgs, size, normalise=False)                 # what cv2.resize itself returns
        ten = data.resize_to_tensor(imgs, size)                                  # + ToTensor
        lab = data.resize_labels(masks, size)
        prod = data.resize_to_tensor(imgs, size, torch.bfloat16, channels_last=True, pad_to=8)
        assert lab.dtype == torch.int64 and ten.dtype == torch.float32
        for i, n in enumerate(names):
            want = torch.from_numpy(z[n + ".linear"]).permute(2, 0, 1).float()
            assert torch.equal(raw[i].cpu(), want), n                            # bit-exact bytes
            assert torch.equal(ten[i].cpu(), torch.from_numpy(z[n + ".tensor"])), n
            assert torch.equal(lab[i].cpu(), torch.from_numpy(z[n + ".nearest"]).long()), n
            assert torch.equal(prod[i, :3].float().cpu(), torch.from_numpy(z[n + ".tensor"]).to(torch.bfloat16).float()), n
        assert prod.is_contiguous(memory_format=torch.channels_last) and float(prod[:, 3:].abs().max()) == 0.0


def test_resize_matches_oracle_on_dataset_sizes():
    """Sizes the goldens are too small for (ADE20K 512x683, Cityscapes 1024x2048, COCO 480x640) against the oracle,
    which tests/test_oracle_golden.py pins to cv2's outputs and tests/test_oracle_vs_reference.py fuzzes against cv2."""
    from maskunet_b200 import data
    from oracle import resize_oracle as ro
    rng = np.random.default_rng(5)
    for sh, sw in ((512, 683), (1024, 2048), (480, 640), (333, 500), (128, 128), (1, 7)):
        img = rng.integers(0, 256, size=(sh, sw, 3), dtype=np.uint8)
        mask = rng.integers(0, 256, size=(sh, sw), dtype=np.uint8)
        got = data.resize_to_tensor([torch.from_numpy(img).to(DEV)], (128, 128), normalise=False)[0]
        want = torch.from_numpy(ro.resize_linear_u8(img, (128, 128))).permute(2, 0, 1).float()
        assert torch.equal(got.cpu(), want), (sh, sw)
        gl = data.resize_labels([torch.from_numpy(mask).to(DEV)], (128, 128))[0]
        assert torch.equal(gl.cpu(), torch.from_numpy(ro.resize_nearest_u8(mask, (128, 128))).long()), (sh, sw)
