"""CPU restatement of the reference's InstanceContrastiveLoss.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/ and the golden generator, never by the product.

Follows /root/reference/code/coco/coco_panoptic.py:482-521 (identical class in ade20k/ade_panoptic.py:390- and
cityscapes/city_panoptic.py:426-) and the 255-ignoring variant cityscapes/city_instance.py:279-307.  Pinned against
the reference's own classes by tests/golden/instance_loss_*.npz (tests/golden/make_golden.py) and live by
tests/test_oracle_vs_reference.py.

What the reference computes, restated without its Python loop over nonzero() results:
  * instances = sorted unique ids of instance_mask [B, H, W], id 0 skipped (:495), 255 skipped in the Cityscapes
    variant (its unique() runs over the valid pixels only, city_instance.py:285-286);
  * an instance with fewer than 2 pixels (:498) or covering every pixel (:507) is skipped;
  * anchor / positive = its first two pixels in row-major order, negative = the k-th pixel NOT in the instance with
    k = torch.randint(0, n_negative, (1,)) from the CPU default generator (:510), one draw per qualifying instance in
    ascending id order;
  * a pixel (b, h, w) selects the logit column sem[:, :, b, h] -- the first TWO components of the 3-component
    nonzero tuple are used as (h, w) (:502-503, :511) -- flattened to a B*C vector;
  * loss = mean over instances of max(||a - p + 1e-6|| - ||a - n + 1e-6|| + margin, 0)  (nn.TripletMarginLoss).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

EPS = 1e-6   # nn.TripletMarginLoss default, added to the difference inside the norm (F.pairwise_distance)


def select_pixels(instance_mask: torch.Tensor, ignore_value: Optional[int] = None) -> List[Tuple[int, int, int, int]]:
    """[(id, p_anchor, p_positive, p_negative)] as row-major positions over [B, H, W]; draws from the CPU generator."""
    flat = instance_mask.reshape(-1).cpu().numpy()
    M = flat.size
    order = np.argsort(flat, kind="stable")
    ids, start, count = np.unique(flat[order], return_index=True, return_counts=True)
    out = []
    for i, s, c in zip(ids.tolist(), start.tolist(), count.tolist()):
        if i == 0 or (ignore_value is not None and i == ignore_value) or c < 2 or c == M:
            continue
        k = int(torch.randint(0, M - c, (1,)))
        members = order[s:s + c]                               # ascending positions of the instance
        gaps = members - np.arange(c)                          # non-members below each member
        t = int(np.searchsorted(gaps, k, side="right"))        # members below the k-th non-member
        out.append((i, int(members[0]), int(members[1]), k + t))
    return out


def instance_contrastive_loss(sem: torch.Tensor, instance_mask: torch.Tensor, margin: float = 1.0,
                              ignore_value: Optional[int] = None):
    """Returns (loss, selections).  `sem` [B, C, H, W] float (autograd flows to it), instance_mask int64 [B, H, W]."""
    B, H, W = instance_mask.shape
    sel = select_pixels(instance_mask, ignore_value)
    if not sel:
        return torch.tensor(0.0), sel
    pos = torch.tensor([[a, p, n] for _, a, p, n in sel], dtype=torch.int64)        # [K, 3]
    i0, i1 = pos // (H * W), (pos // W) % H                                           # (batch, row) used as (h, w)
    cols = sem.float()[:, :, i0, i1]                                                  # [B, C, K, 3]
    cols = cols.permute(2, 3, 0, 1).reshape(len(sel), 3, -1)                          # [K, 3, B*C]
    d_ap = (cols[:, 0] - cols[:, 1] + EPS).square().sum(-1).sqrt()
    d_an = (cols[:, 0] - cols[:, 2] + EPS).square().sum(-1).sqrt()
    return torch.clamp_min(d_ap - d_an + margin, 0).mean(), sel
