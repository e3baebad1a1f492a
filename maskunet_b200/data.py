"""Device half of the input pipeline.

The scripts build their batches on the main thread with cv2 and ``torchvision.transforms.ToTensor()``
(/root/reference/code/ade20k/ade_semantic.py:56-79, 85, 97: ``num_workers=0``) and copy fp32 NCHW tensors to the GPU.
Here the batch crosses PCIe as uint8 HWC (4x fewer bytes) and ``to_tensor`` -- ``mu_to_tensor_u8`` -- writes the
network input on the device, bit-exact with ToTensor (IEEE ``/ 255``), optionally straight into the production layout
(bf16 channels-last, channels zero-padded to the 8 the tcgen05 stem convolution reads).  ``DevicePrefetcher`` keeps one
batch in flight on a side stream from pinned buffers.

The resize of the dataset classes (``cv2.resize(image, (128, 128), INTER_LINEAR)`` / ``cv2.resize(mask, ..., INTER_NEAREST)``,
:72-73) runs on the device too: ``resize_to_tensor`` / ``resize_labels`` take the decoded images at their ORIGINAL sizes
(uint8, one tensor per image: sizes differ) and write the network input / the int64 labels directly, bit-exact with
OpenCV's uint8 fixed-point arithmetic (``mu_resize_linear_to_tensor_u8``, ``mu_resize_nearest_u8_i64``; oracle
``oracle/resize_oracle.py`` pinned against cv2 itself).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Iterator, Sequence, Tuple

import torch
from torch import Tensor

from . import ops
from ._lib import check

_L = ops._L


@torch.library.custom_op("maskunet::to_tensor_u8", mutates_args=(), device_types="cuda")
def to_tensor_u8(images: Tensor, bf16: bool, channels_last: bool, pad_to: int) -> Tensor:
    """uint8 [B, H, W, C] -> [B, max(C, pad_to), H, W] = images / 255 (channels >= C zero), NCHW or channels-last memory."""
    if not images.is_cuda or images.dtype != torch.uint8 or images.dim() != 4 or not images.is_contiguous():
        raise RuntimeError("to_tensor: expects a contiguous CUDA uint8 batch [B, H, W, C] (there is no CPU fallback)")
    B, H, W, C = images.shape
    cpad = max(C, pad_to) if channels_last else C
    dtype = torch.bfloat16 if bf16 else torch.float32
    out = torch.empty((B, cpad, H, W), dtype=dtype, device=images.device,
                      memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    with torch.cuda.device(images.device):
        ops._count(1)
        check(_L.mu_to_tensor_u8(ops._p(images), ops._p(out), B, H, W, C, cpad, int(channels_last), ops._code(out),
                                 ops._stream(images)), "mu_to_tensor_u8")
    return out


@to_tensor_u8.register_fake
def _(images, bf16, channels_last, pad_to):
    B, H, W, C = images.shape
    cpad = max(C, pad_to) if channels_last else C
    return torch.empty((B, cpad, H, W), dtype=torch.bfloat16 if bf16 else torch.float32, device=images.device,
                       memory_format=torch.channels_last if channels_last else torch.contiguous_format)


def to_tensor(images: Tensor, dtype: torch.dtype = torch.float32, channels_last: bool = False, pad_to: int = 0) -> Tensor:
    """Batched ``ToTensor()`` on the device.  Defaults give the reference's fp32 NCHW tensor; the production
    configuration is ``to_tensor(u8, torch.bfloat16, channels_last=True, pad_to=8)``."""
    if not images.is_cuda:
        raise RuntimeError("to_tensor: expects a CUDA uint8 batch [B, H, W, C] (there is no CPU fallback)")
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("to_tensor: dtype must be float32 or bfloat16")
    if pad_to and not channels_last:
        raise ValueError("to_tensor: channel padding needs channels_last=True")
    return to_tensor_u8(images, dtype == torch.bfloat16, channels_last, pad_to)


def resize_to_tensor(images: Sequence[Tensor], size: Tuple[int, int] = (128, 128), dtype: torch.dtype = torch.float32,
                     channels_last: bool = False, pad_to: int = 0, normalise: bool = True) -> Tensor:
    """``ToTensor()(cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR))`` for every image of a batch, on the device.

    images: CUDA uint8 tensors [h_i, w_i, C] (HWC as cv2 delivers them; the sizes may differ).  size = (H, W).
    Returns [B, max(C, pad_to), H, W] (NCHW or channels-last memory).  ``normalise=False`` keeps the resized bytes as
    numbers 0..255 (what cv2.resize itself returns) instead of dividing by 255."""
    if not images:
        raise ValueError("resize_to_tensor: empty batch")
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("resize_to_tensor: dtype must be float32 or bfloat16")
    if pad_to and not channels_last:
        raise ValueError("resize_to_tensor: channel padding needs channels_last=True")
    oh, ow = size
    C = images[0].shape[2]
    cpad = max(C, pad_to) if channels_last else C
    dev = images[0].device
    out = torch.empty((len(images), cpad, oh, ow), dtype=dtype, device=dev,
                      memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    code = ops._code(out)
    slot = cpad * oh * ow * out.element_size()
    with torch.cuda.device(dev):
        for i, img in enumerate(images):
            if not (img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
                    and img.shape[2] == C and img.device == dev):
                raise RuntimeError("resize_to_tensor: expects contiguous CUDA uint8 images [h, w, C] (no CPU fallback)")
            ops._count(1)
            check(_L.mu_resize_linear_to_tensor_u8(ops._p(img), ctypes.c_void_p(out.data_ptr() + i * slot), img.shape[0],
                                                   img.shape[1], C, oh, ow, cpad, int(channels_last), int(normalise),
                                                   code, ops._stream(out)), "mu_resize_linear_to_tensor_u8")
    return out


def resize_labels(masks: Sequence[Tensor], size: Tuple[int, int] = (128, 128)) -> Tensor:
    """``torch.from_numpy(cv2.resize(mask, (W, H), interpolation=cv2.INTER_NEAREST)).long()`` (:73, :78) for every label
    map of a batch, on the device.  masks: CUDA uint8 tensors [h_i, w_i].  Returns int64 [B, H, W]."""
    if not masks:
        raise ValueError("resize_labels: empty batch")
    oh, ow = size
    dev = masks[0].device
    out = torch.empty((len(masks), oh, ow), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        for i, m in enumerate(masks):
            if not (m.is_cuda and m.dtype == torch.uint8 and m.dim() == 2 and m.is_contiguous() and m.device == dev):
                raise RuntimeError("resize_labels: expects contiguous CUDA uint8 label maps [h, w] (no CPU fallback)")
            ops._count(1)
            check(_L.mu_resize_nearest_u8_i64(ops._p(m), ctypes.c_void_p(out.data_ptr() + i * oh * ow * 8), m.shape[0],
                                              m.shape[1], oh, ow, ops._stream(out)), "mu_resize_nearest_u8_i64")
    return out


class DevicePrefetcher:
    """Iterate ``(uint8 HWC images, *labels)`` host batches one step ahead: pinned staging, copies and the ToTensor
    kernel on a side stream, the consumer's stream waits on an event (no host sync)."""

    def __init__(self, batches: Iterable[Tuple[Tensor, ...]], device: torch.device, dtype=torch.bfloat16,
                 channels_last: bool = True, pad_to: int = 8):
        self.batches, self.device = batches, device
        self.kw = dict(dtype=dtype, channels_last=channels_last, pad_to=pad_to if channels_last else 0)
        self.stream = torch.cuda.Stream(device)

    def _stage(self, batch):
        with torch.cuda.stream(self.stream):
            dev = [t if t.is_cuda else (t if t.is_pinned() else t.pin_memory()).to(self.device, non_blocking=True)
                   for t in batch]
            out = (to_tensor(dev[0].contiguous(), **self.kw), *dev[1:])
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self) -> Iterator[Tuple[Tensor, ...]]:
        it = iter(self.batches)
        nxt = None
        for batch in it:
            cur, nxt = nxt, self._stage(batch)
            if cur is not None:
                yield self._hand_over(cur)
        if nxt is not None:
            yield self._hand_over(nxt)

    def _hand_over(self, staged):
        out, ev = staged
        torch.cuda.current_stream(self.device).wait_event(ev)
        for t in out:
            t.record_stream(torch.cuda.current_stream(self.device))
        return out
