"""Data-parallel gradient exchange: one process per GPU, bucketed all-reduce overlapped with backward.

The reference's only multi-GPU mechanism is single-process ``torch.nn.DataParallel``
(/root/reference/code/ade20k/ade_semantic.py:373; nine call sites), which scatters
the batch, replicates the module every forward and reduces gradients onto GPU 0.  The B200 design keeps
its arithmetic (per-replica BatchNorm statistics, per-replica masks, mean gradient over the global
batch) but runs one process per GPU: every rank owns batch/world samples and the only exchange step is a
sum of gradients over NCCL/NVLink, issued bucket by bucket from post-accumulate-grad hooks so that it
overlaps the rest of backward.

Parameters that never receive a gradient (the reference's dead ``emb_layer.*``)
are discovered on the first step and left out of the buckets, otherwise the reducer would wait forever.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("params", "pending", "flat", "work")

    def __init__(self, params):
        self.params: List[torch.nn.Parameter] = params
        self.pending = len(params)
        self.flat: Optional[torch.Tensor] = None
        self.work = None


class GradReducer:
    """Bucketed asynchronous all-reduce(mean) of parameter gradients.

    Usage per step:  ``loss.backward(); reducer.finish(); optimizer.step()``.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 * 1024 * 1024,
                 process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: List[_Bucket] = []
        self._bucket_of = {}
        self._order: List[torch.nn.Parameter] = []   # hook firing order during the discovery step
        self._discovered = False
        self.launched = 0                            # all-reduce launches, for reporting
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    # -- hooks ---------------------------------------------------------------------------------------
    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self._discovered:
            self._order.append(p)
            return
        b = self._bucket_of.get(id(p))
        if b is None:   # a parameter that had no gradient during discovery now has one
            raise RuntimeError("GradReducer: parameter produced a gradient after bucket discovery")
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket) -> None:
        b.flat = torch.cat([p.grad.reshape(-1) for p in b.params])
        b.flat.div_(self.world)
        b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.launched += 1

    def _build_buckets(self) -> None:
        cur, size = [], 0
        for p in self._order:   # reverse-forward order: the order gradients become ready
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= self.bucket_bytes:
                self.buckets.append(_Bucket(cur))
                cur, size = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        for b in self.buckets:
            for p in b.params:
                self._bucket_of[id(p)] = b
        self._discovered = True

    # -- per step ------------------------------------------------------------------------------------
    def finish(self) -> None:
        """Wait for every bucket and write the averaged gradients back.  Call after backward."""
        if self.world == 1:
            return
        if not self._discovered:        # first step: learn which parameters get gradients, reduce synchronously
            self._build_buckets()
            for b in self.buckets:
                b.pending = 0
                self._launch(b)
        for b in self.buckets:
            if b.pending != 0:
                raise RuntimeError("GradReducer.finish(): a bucket is missing gradients "
                                   f"({b.pending} of {len(b.params)} parameters did not report)")
            b.work.wait()
            offset = 0
            views = []
            for p in b.params:
                n = p.numel()
                views.append(b.flat[offset:offset + n].view_as(p.grad))
                offset += n
            torch._foreach_copy_([p.grad for p in b.params], views)
            b.flat, b.work, b.pending = None, None, len(b.params)

    def broadcast_parameters(self, module: torch.nn.Module, src: int = 0) -> None:
        """One-time: rank ``src``'s parameters and buffers to every rank (what DataParallel's per-forward
        replicate achieves)."""
        if self.world == 1:
            return
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)
