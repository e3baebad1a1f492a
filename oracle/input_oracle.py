"""CPU restatement of the reference's image-to-tensor step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The scripts apply ``torchvision.transforms.ToTensor()`` to the RGB uint8 HWC array cv2 produced
(/root/reference/code/ade20k/ade_semantic.py:9, 56-79, 85).  torchvision is a third-party dependency of the reference
(``torchvision==0.14.1``, requirement.txt:290), not vendored; its published arithmetic for a uint8 ndarray is
``torch.from_numpy(pic.transpose(2, 0, 1)).to(float32).div(255)``.  Pinned against torchvision's own ToTensor
(0.26 in the build container) by tests/golden/to_tensor.npz (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def to_tensor(images_u8: np.ndarray) -> np.ndarray:
    """uint8 [B, H, W, C] -> float32 [B, C, H, W] = value / 255 (one correctly rounded IEEE division per value)."""
    return (images_u8.astype(np.float32) / np.float32(255)).transpose(0, 3, 1, 2)
