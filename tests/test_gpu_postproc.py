"""GPU parity of the logit post-processing kernels against the golden outputs of the
reference's own mean_iou (ade_semantic.py:128-146) and against the CPU oracle: class maps bit-exact, mean IoU to 1e-6."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("name,C", [("postproc_c150", 150), ("postproc_c19", 19)])
@pytest.mark.parametrize("layout", ["bf16_padded_view", "bf16_channels_last", "fp32_nchw"])
def test_class_map_and_mean_iou_match_reference_golden(name, C, layout):
    from maskunet_b200 import ops
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    logits, labels = torch.from_numpy(z["logits"]).to(DEV), torch.from_numpy(z["labels"]).to(DEV)
    if layout == "bf16_padded_view":            # what the 1x1 head returns: first C channels of a class-padded buffer
        P = ops.pad_channels(C)
        buf = torch.zeros(logits.shape[0], P, *logits.shape[2:], device=DEV, dtype=torch.bfloat16)
        buf = buf.contiguous(memory_format=torch.channels_last)
        buf[:, :C] = logits.to(torch.bfloat16)
        y = buf[:, :C]
    elif layout == "bf16_channels_last":
        y = logits.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    else:
        y = logits
    pred = ops.segmentation_argmax(y)
    assert pred.dtype == torch.int64
    assert torch.equal(pred.cpu(), torch.from_numpy(z["pred"].astype(np.int64)))          # identical class maps
    miou = ops.mean_iou(y, labels, C)
    assert abs(float(miou) - float(z["miou"][0])) < 1e-6


def test_mean_iou_full_size_properties():
    """BASELINE size (256 x 150 x 128 x 128): histogram identities and agreement with the oracle on a sample."""
    from maskunet_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(0)
    B, C = 64, 150
    y = torch.relu(torch.randn(B, C, 128, 128, device=DEV, generator=g)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    labels = torch.randint(0, C, (B, 128, 128), device=DEV, generator=g)
    labels[:, :2] = 255
    t, pitch = ops._class_rows(y)
    pred, miou, hist = ops.argmax_iou(t, labels, pitch, 1e-6)
    assert int(hist[0].sum()) == B * 128 * 128                       # every pixel is predicted as exactly one class
    assert int(hist[1].sum()) == int((labels != 255).sum())
    assert bool((hist[2] <= torch.minimum(hist[0], hist[1])).all())
    assert torch.equal(pred[:2].cpu(), mo.class_map(y[:2].float().cpu()))
    ref = mo.mean_iou(y.float().cpu(), labels.cpu(), C)
    assert abs(float(miou) - float(ref)) < 1e-6
    # all-equal logits: ties go to the lowest index
    assert int(ops.segmentation_argmax(torch.zeros(1, C, 16, 16, device=DEV, dtype=torch.bfloat16).contiguous(
        memory_format=torch.channels_last)).max()) == 0
