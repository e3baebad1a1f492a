"""InstanceContrastiveLoss on the device.

Drop-in for the class every panoptic / instance script defines inline
(/root/reference/code/coco/coco_panoptic.py:482-521; ade20k/ade_panoptic.py:390-; cityscapes/city_panoptic.py:426-;
cityscapes/city_instance.py:279-307 is the same with ``ignore_value=255``): same constructor ``(margin=1.0)``, same
``forward(sem_mask, instance_mask)``, same value, same consumption of the CPU random generator (one
``torch.randint(0, n_negative, (1,))`` per qualifying instance, ascending id order, :510).

The reference loops in Python over ``torch.unique`` with two ``.nonzero()`` host syncs per instance.  Here:
  1. the ids are grouped on the device (one stable sort) and ONE small device->host copy brings back the distinct
     ids and their pixel counts -- the host needs the counts, they are the bounds of the randint draws;
  2. ``mu_instance_triplet_fwd`` selects anchor / positive / k-th non-member pixel per instance (binary search in the
     sorted group), gathers the three logit columns and accumulates the mean triplet loss; ``..._bwd`` scatters the
     gradient.  ``accumulate_grad`` adds it straight into an existing gradient buffer (the fused cross-entropy's).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import ops
from ._lib import check

_L = ops._L
TRIPLET_EPS = 1e-6          # nn.TripletMarginLoss default


def _strides(t: Tensor):
    return (ctypes.c_int64 * 4)(*t.stride())


def plan_instances(instance_mask: Tensor, ignore_value: Optional[int] = None) -> Tuple[Tensor, Tensor, int]:
    """(order int64 [M], meta int64 [3, K] on the device, K).  One host sync; K draws from the CPU generator."""
    if not instance_mask.is_cuda or instance_mask.dim() != 3:
        raise RuntimeError("InstanceContrastiveLoss: instance_mask must be a CUDA tensor [B, H, W] (no CPU fallback)")
    flat = instance_mask.reshape(-1).to(torch.int64)
    M = flat.numel()
    sorted_ids, order = torch.sort(flat, stable=True)
    ids, counts = torch.unique_consecutive(sorted_ids, return_counts=True)
    host = torch.stack([ids, counts]).cpu()                      # the one device->host copy of the loss
    ids_h, counts_h = host[0].tolist(), host[1].tolist()
    off, rows = 0, []
    for i, c in zip(ids_h, counts_h):
        if not (i == 0 or (ignore_value is not None and i == ignore_value) or c < 2 or c == M):
            rows.append((off, c, int(torch.randint(0, M - c, (1,)))))     # :510, CPU default generator
        off += c
    K = len(rows)
    if K == 0:
        return order, torch.empty((3, 0), dtype=torch.int64, device=flat.device), 0
    meta = torch.tensor(rows, dtype=torch.int64).t().contiguous().pin_memory().to(flat.device, non_blocking=True)
    return order, meta, K


@torch.library.custom_op("maskunet::instance_triplet", mutates_args=(), device_types="cuda")
def instance_triplet(sem: Tensor, order: Tensor, meta: Tensor, margin: float) -> Tuple[Tensor, Tensor, Tensor]:
    """sem [B, C, H, W] (any strides), order / meta from plan_instances -> (loss f32 [1], sel int32 [K, 6], dist f32 [K, 2])."""
    B, C, H, W = sem.shape
    K = meta.shape[1]
    dev = sem.device
    loss = torch.empty((1,), dtype=torch.float32, device=dev)
    sel = torch.empty((K, 6), dtype=torch.int32, device=dev)
    dist = torch.empty((K, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ops._count(2 if K else 0)
        check(_L.mu_instance_triplet_fwd(ops._p(sem), _strides(sem), B, C, H, W, ops._p(order), ops._p(meta), K,
                                         margin, TRIPLET_EPS, ops._p(sel), ops._p(dist), ops._p(loss), ops._code(sem),
                                         ops._stream(sem)), "mu_instance_triplet_fwd")
    return loss, sel, dist


@instance_triplet.register_fake
def _(sem, order, meta, margin):
    K = meta.shape[1]
    return (sem.new_empty((1,), dtype=torch.float32), sem.new_empty((K, 6), dtype=torch.int32),
            sem.new_empty((K, 2), dtype=torch.float32))


def accumulate_grad(sem: Tensor, sel: Tensor, dist: Tensor, margin: float, dsem: Tensor, scale: float = 1.0,
                    dloss: Optional[Tensor] = None) -> None:
    """dsem += scale * dloss * d(loss)/d(sem), in place; dsem is [B, C, H, W] with any strides, dtype of sem."""
    B, C, H, W = sem.shape
    assert dsem.shape == sem.shape and dsem.dtype == sem.dtype and dsem.device == sem.device
    K = sel.shape[0]
    with torch.cuda.device(sem.device):
        ops._count(1 if K else 0)
        check(_L.mu_instance_triplet_bwd(ops._p(sem), _strides(sem), B, C, ops._p(sel), K, margin, TRIPLET_EPS,
                                         ops._p(dist), ops._optp(dloss), scale, ops._p(dsem), _strides(dsem),
                                         ops._code(sem), ops._stream(sem)), "mu_instance_triplet_bwd")


@torch.library.custom_op("maskunet::instance_triplet_bwd", mutates_args=(), device_types="cuda")
def instance_triplet_bwd(sem: Tensor, sel: Tensor, dist: Tensor, dloss: Tensor, margin: float) -> Tensor:
    dsem = torch.zeros_like(sem)                      # preserves the (dense) stride order of sem
    accumulate_grad(sem, sel, dist, margin, dsem, 1.0, dloss.to(torch.float32).reshape(1).contiguous())
    return dsem


@instance_triplet_bwd.register_fake
def _(sem, sel, dist, dloss, margin):
    return torch.empty_like(sem)


def _it_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    sem, order, meta, margin = inputs
    ctx.margin = margin
    ctx.save_for_backward(sem, output[1], output[2])


def _it_backward(ctx, dloss, *unused):
    sem, sel, dist = ctx.saved_tensors
    return instance_triplet_bwd(sem, sel, dist, dloss, ctx.margin), None, None, None


instance_triplet.register_autograd(_it_backward, setup_context=_it_setup)


class InstanceContrastiveLoss(nn.Module):
    """``InstanceContrastiveLoss(margin=1.0)``; ``forward(sem_mask [B, C, H, W], instance_mask int [B, H, W])``.

    ``ignore_value`` (keyword-only, default None = the COCO / ADE20K class) excludes that id like the background id 0:
    255 reproduces cityscapes/city_instance.py:285-286.  As in the reference, a pixel's (batch, row) index pair is used
    as the (h, w) of the logit column (:502-503), which needs B <= H; where the reference raises IndexError the loss
    here is NaN (the check happens on the device, without a host sync)."""

    def __init__(self, margin: float = 1.0, *, ignore_value: Optional[int] = None):
        super().__init__()
        self.margin = margin
        self.ignore_value = ignore_value

    def forward(self, sem_mask: Tensor, instance_mask: Tensor) -> Tensor:
        if not sem_mask.is_cuda:
            raise RuntimeError("InstanceContrastiveLoss runs on CUDA tensors only (there is no CPU fallback)")
        if sem_mask.dim() != 4 or tuple(instance_mask.shape) != (sem_mask.shape[0], *sem_mask.shape[2:]):
            raise ValueError("InstanceContrastiveLoss: expected sem_mask [B, C, H, W] and instance_mask [B, H, W]")
        order, meta, K = plan_instances(instance_mask, self.ignore_value)
        if K == 0:
            return torch.tensor(0.0, device=sem_mask.device)          # :521
        return instance_triplet(sem_mask, order, meta, float(self.margin))[0].squeeze(0)
