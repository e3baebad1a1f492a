"""Data-parallel gradient exchange: one process per GPU, bucketed all-reduce overlapped with backward.

The reference's only multi-GPU mechanism is single-process ``torch.nn.DataParallel``
(/root/reference/code/ade20k/ade_semantic.py:373; nine call sites), which scatters
the batch, replicates the module every forward and reduces gradients onto GPU 0.  The B200 design keeps
its arithmetic (per-replica BatchNorm statistics, per-replica masks, mean gradient over the global
batch) but runs one process per GPU: every rank owns batch/world samples and the only exchange step is a
mean of gradients over NCCL/NVLink, issued bucket by bucket from post-accumulate-grad hooks so that it
overlaps the rest of backward.

Buckets are persistent flat fp32 buffers.  When the last gradient of a bucket is ready, ONE multi-tensor
copy packs the bucket, ``all_reduce(AVG)`` runs in place on NCCL's stream, and ``finish()`` only waits and
re-points every ``p.grad`` at its slice of the flat buffer -- no ``cat``, no divide, no copy back.

Parameters that never receive a gradient (the reference's dead ``emb_layer.*``)
are discovered on the first step and left out of the buckets, otherwise the reducer would wait forever.
``step()`` of the Trainer returns the rank-local loss share; gradients, not losses, are what is exchanged.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("params", "pending", "flat", "views", "work", "seen")

    def __init__(self, params, device, dtype):
        self.params: List[torch.nn.Parameter] = params
        self.pending = len(params)
        self.flat = torch.zeros(sum(p.numel() for p in params), dtype=dtype, device=device)
        self.views, off = [], 0
        for p in params:
            # the slice carries the parameter's own strides (channels-last convolution weights are dense but permuted):
            # the fused optimiser insists on gradients with the parameter's layout
            dense = p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last)
            self.views.append(self.flat.as_strided(p.size(), p.stride(), off) if dense
                              else self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.work = None
        self.seen = set()


class GradReducer:
    """Bucketed asynchronous all-reduce(mean) of parameter gradients.

    Usage per step:  ``loss.backward(); reducer.finish(); optimizer.step()``.  With gradient accumulation call
    ``reducer.accumulate(True)`` before every backward but the last and ``accumulate(False)`` before the last: the
    hooks then stay silent until the final backward, whose gradients (the accumulated sums) are exchanged.
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
        self._order_ids = set()
        self._discovered = False
        self._accumulating = False
        self.launched = 0                            # all-reduce launches, for reporting
        backend = dist.get_backend(process_group) if dist.is_initialized() else ""
        self._avg = backend == "nccl"                # ReduceOp.AVG exists on NCCL only (gloo: sum, then divide)
        # measurement switch (bench.py --ddp-dryrun): hooks and bucket packing run, the collective does not -- the
        # difference to a normal run is what the exchange itself (NCCL kernels + rank synchronisation) costs
        self.dry_run = False
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def accumulate(self, on: bool) -> None:
        """on=True: the next backward only accumulates into .grad (no exchange)."""
        self._accumulating = bool(on)

    # -- hooks ---------------------------------------------------------------------------------------
    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self._accumulating:
            return
        if not self._discovered:
            if id(p) not in self._order_ids:         # a parameter used twice fires once per accumulation
                self._order_ids.add(id(p))
                self._order.append(p)
            return
        b = self._bucket_of.get(id(p))
        if b is None:   # a parameter that had no gradient during discovery now has one
            raise RuntimeError("GradReducer: parameter produced a gradient after bucket discovery")
        if id(p) in b.seen:
            return
        b.seen.add(id(p))
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket) -> None:
        grads = [p.grad for p in b.params]
        if any(g.data_ptr() != v.data_ptr() for g, v in zip(grads, b.views)):
            torch._foreach_copy_(b.views, grads)     # one multi-tensor launch packs the bucket
        for p, v in zip(b.params, b.views):
            p.grad = v                               # frees autograd's buffer; the optimiser reads the reduced slice
        if self.dry_run:
            b.work = None
        elif self._avg:
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            b.flat.div_(self.world)
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.launched += 1

    def _build_buckets(self) -> None:
        if not self._order:
            raise RuntimeError("GradReducer.finish() called before any backward produced a gradient")
        cur, size = [], 0
        def close():
            self.buckets.append(_Bucket(cur, cur[0].device, cur[0].dtype))
        for p in self._order:   # reverse-forward order: the order gradients become ready
            if cur and p.dtype != cur[0].dtype:
                close()
                cur, size = [], 0
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= self.bucket_bytes:
                close()
                cur, size = [], 0
        if cur:
            close()
        for b in self.buckets:
            for p in b.params:
                self._bucket_of[id(p)] = b
        self._discovered = True

    # -- per step ------------------------------------------------------------------------------------
    def finish(self) -> None:
        """Wait for every bucket; afterwards every ``p.grad`` is the mean gradient.  Call after backward."""
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
            if b.work is not None:
                b.work.wait()
            b.work, b.pending = None, len(b.params)
            b.seen.clear()

    def broadcast_parameters(self, module: torch.nn.Module, src: int = 0) -> None:
        """One-time: rank ``src``'s parameters and buffers to every rank (what DataParallel's per-forward
        replicate achieves)."""
        if self.world == 1:
            return
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)
