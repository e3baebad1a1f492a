"""The reference's timed loop body as a reusable step.

/root/reference/code/ade20k/ade_semantic.py:394-401:
    inputs, labels = inputs.to(device), labels.to(device)
    optimizer.zero_grad(); outputs = model(inputs); loss = criterion(outputs, labels)
    loss.backward(); optimizer.step()
with ``criterion = nn.CrossEntropyLoss()`` and ``optim.AdamW(lr=5e-5, weight_decay=1e-1)`` (:377-379).
Under data parallelism the gradient all-reduce (ddp.GradReducer) sits between backward and the optimiser.

The panoptic scripts add the instance term (coco/coco_panoptic.py:544-553):
    loss = 0.9 * semantic_loss_fn(logits, semantic_labels) + 0.1 * instance_loss_fn(logits, instance_labels)
-- pass ``instance_loss=InstanceContrastiveLoss()`` and give ``step`` the instance labels.

Loss normalisation.  The reference computes ONE mean over the whole (global) batch: ``nn.DataParallel`` gathers the
outputs on GPU 0 before the criterion (:373, :399).  Here every rank -- and, with ``micro_batch``, every slice of a
rank's batch -- normalises its cross-entropy sum by the number of valid pixels of the WHOLE global batch divided by the
world size, so that the all-reduce(mean) of the gradients is exactly the gradient of that global mean even when
``ignore_index`` pixels are spread unevenly.  ``step`` returns the rank-local share of the loss (device tensor, no
host sync); its mean over ranks is the global loss.

CUDA graph (``Trainer(cuda_graph=True)``).  A step is ~460 kernel launches of ours plus torch's own; enqueuing them
costs the host ~25 ms, which is hidden behind 115 ms of GPU work at 256 images per GPU but not at the small per-GPU
batches strong scaling produces.  With the flag, the first ``graph_warmup`` steps run eagerly (masks are drawn, kernel
attributes set, TMA descriptors cached, optimiser state created), then zero_grad + forward + fused loss + backward +
AdamW(capturable) are captured ONCE into a CUDA graph with static input buffers; every later step is two copies into
those buffers and one graph launch.  Shapes must not change afterwards.  The instance term needs a host round trip per
step (the reference's ``torch.randint`` draws from the CPU generator, coco_panoptic.py:510) and micro-batching changes
the step's structure, so both stay eager.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .ddp import GradReducer


class Trainer:
    def __init__(self, model: torch.nn.Module, lr: float = 5e-5, weight_decay: float = 1e-1,
                 ignore_index: int = -100, data_parallel: bool = False, bucket_bytes: int = 25 * 1024 * 1024,
                 instance_loss: Optional[torch.nn.Module] = None, loss_weights=(0.9, 0.1),
                 cuda_graph: bool = False, graph_warmup: int = 3):
        self.model = model
        self.device = next(model.parameters()).device
        self.ignore_index = ignore_index
        self.instance_loss = instance_loss          # maskunet_b200.InstanceContrastiveLoss or None
        self.loss_weights = loss_weights            # (semantic, instance), coco_panoptic.py:552
        params = [p for p in model.parameters() if p.requires_grad]
        self.cuda_graph = bool(cuda_graph) and self.device.type == "cuda"
        self.graph_warmup = max(1, int(graph_warmup))
        self._graph = None               # (CUDAGraph, static images, static labels, static loss)
        self._eager_steps = 0
        self.graph_launches = 0          # kernels of ours inside one replay (ops.LAUNCHES only sees the capture)
        self.optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, fused=self.device.type == "cuda",
                                           capturable=self.cuda_graph)
        self.reducer: Optional[GradReducer] = None
        if data_parallel:
            self.reducer = GradReducer(params, bucket_bytes=bucket_bytes)
            self.reducer.broadcast_parameters(model)

    # ------------------------------------------------------------------------------------------ loss pieces
    def _valid_count(self, labels: torch.Tensor, sliced: bool) -> Optional[torch.Tensor]:
        """f32 [1] normaliser of the cross-entropy sum: valid pixels of the global batch / world size.  None when one
        pass over this rank's batch with its own count is already that (the fused kernel then counts itself)."""
        world = self.reducer.world if self.reducer is not None else 1
        uneven = world > 1 and self.ignore_index >= 0          # ranks may hold different numbers of valid pixels
        if not (sliced or uneven):
            return None
        count = (labels != self.ignore_index).sum().to(torch.float32).reshape(1)
        if uneven:
            dist.all_reduce(count, group=self.reducer.group)   # 4 bytes
            count = count / world
        return count

    def _drop_graph_refs(self) -> None:
        """The model keeps the class-padded logits of its last forward (the tensor backward starts from).  After
        backward nothing needs it, and holding it would keep that iteration's autograd graph -- including the
        parameters' AccumulateGrad nodes, which are bound to the stream they were created on -- alive into the next
        iteration: a CUDA-graph capture on another stream then trips over the stale nodes."""
        if hasattr(self.model, "_padded_logits"):
            self.model._padded_logits = None

    def _ce_backward(self, logits, labels, count):
        """Cross-entropy of one (micro-)batch, normalised by ``count``; starts backward; returns the loss share."""
        fused_ok = logits.is_cuda and logits.shape[1] <= 256 and logits.dtype in (torch.float32, torch.bfloat16)
        padded = getattr(self.model, "_padded_logits", None)
        if (fused_ok and padded is not None and padded.requires_grad and padded.data_ptr() == logits.data_ptr()
                and padded.is_contiguous(memory_format=torch.channels_last)):
            # logits are the first c_out channels of the head's class-padded buffer: run the fused loss on the
            # buffer itself and start backward there (no narrow / pad passes, no loss node)
            with torch.no_grad():
                loss, dpad = ops.cross_entropy_fused(padded.detach(), labels, self.ignore_index, logits.shape[1], count)
            padded.backward(dpad)
            self._drop_graph_refs()
            return loss.squeeze(0)
        if fused_ok and logits.is_contiguous(memory_format=torch.channels_last):
            # fused CrossEntropyLoss(mean) forward + gradient on channels-last logits (one read, one write), and
            # backward starts from d(loss)/d(logits) directly: the loss node and its `dlogits * dloss` pass are skipped
            with torch.no_grad():
                loss, dlogits = ops.cross_entropy_fused(logits.detach(), labels, self.ignore_index, -1, count)
            logits.backward(dlogits)
            return loss.squeeze(0)
        loss = self._ce_torch(logits, labels, count)
        loss.backward()
        return loss.detach()

    def _ce_torch(self, logits, labels, count):
        if count is None:
            return F.cross_entropy(logits.float(), labels, ignore_index=self.ignore_index)
        return F.cross_entropy(logits.float(), labels, ignore_index=self.ignore_index, reduction="sum") / count[0]

    def _panoptic_backward(self, logits, labels, inst, count, inst_scale, plan=None):
        from . import losses
        w_sem, w_inst = self.loss_weights
        crit = self.instance_loss
        padded = getattr(self.model, "_padded_logits", None)
        fused = (logits.is_cuda and isinstance(crit, losses.InstanceContrastiveLoss) and padded is not None
                 and padded.requires_grad and padded.data_ptr() == logits.data_ptr()
                 and padded.is_contiguous(memory_format=torch.channels_last) and logits.shape[1] <= 256)
        if not fused:
            loss = w_sem * self._ce_torch(logits, labels, count) + (w_inst * inst_scale) * crit(logits, inst)
            loss.backward()
            return loss.detach()
        loss, dpad = self.fused_panoptic_loss(padded, logits.shape[1], labels, inst, count, inst_scale, plan)
        padded.backward(dpad)
        self._drop_graph_refs()
        return loss

    def fused_panoptic_loss(self, padded, c_out, labels, inst, count=None, inst_scale: float = 1.0, plan=None):
        """(w_sem * CE + w_inst * inst_scale * InstanceContrastiveLoss, its gradient w.r.t. the class-padded logit
        buffer).  Both gradients land in ONE buffer: the fused cross-entropy writes d(CE)/d(logits) into it, the
        triplet kernel adds its three logit columns per instance; backward then starts from the sum."""
        from . import losses
        w_sem, w_inst = self.loss_weights
        w_inst = w_inst * inst_scale
        crit = self.instance_loss
        with torch.no_grad():
            padded = padded.detach()
            loss, dpad = ops.cross_entropy_fused(padded, labels, self.ignore_index, c_out, count)
            dpad.mul_(w_sem)
            loss = w_sem * loss.squeeze(0)
            order, meta, K = plan if plan is not None else losses.plan_instances(inst, crit.ignore_value)
            if K:
                view = padded[:, :c_out]
                l_inst, sel, dist_ = losses.instance_triplet(view, order, meta, float(crit.margin))
                losses.accumulate_grad(view, sel, dist_, float(crit.margin), dpad[:, :c_out], scale=w_inst)
                loss = loss + w_inst * l_inst.squeeze(0)
        return loss, dpad

    # ------------------------------------------------------------------------------------------ the step
    def forward_backward(self, images: torch.Tensor, labels: torch.Tensor,
                         instance_labels: Optional[torch.Tensor] = None,
                         micro_batch: Optional[int] = None) -> torch.Tensor:
        """Forward + loss + backward (gradients ACCUMULATE into ``.grad``; no zero_grad, no optimiser).  With
        ``micro_batch`` the batch is processed in slices of that many samples -- the same gradient as one pass for the
        cross-entropy term (every slice is normalised by the valid-pixel count of the whole batch); BatchNorm batch
        statistics and the instance term are per slice, as they are per replica under the reference's DataParallel.
        Returns the loss (device tensor)."""
        if images.device != self.device:
            images = images.to(self.device, non_blocking=True)
        if labels.device != self.device:
            labels = labels.to(self.device, non_blocking=True)
        B = images.shape[0]
        mb = B if not micro_batch or micro_batch >= B else int(micro_batch)
        n_slices = (B + mb - 1) // mb
        count = self._valid_count(labels, n_slices > 1)
        panoptic = self.instance_loss is not None and instance_labels is not None
        plans = None
        if panoptic:
            from . import losses
            if instance_labels.device != self.device:
                instance_labels = instance_labels.to(self.device, non_blocking=True)
            if isinstance(self.instance_loss, losses.InstanceContrastiveLoss):
                # the grouping of the instance ids (one sort + ONE device->host copy per slice, the bounds of the
                # reference's randint draws, coco_panoptic.py:510) depends on the labels only: done for every slice
                # BEFORE the first forward is enqueued, so no host sync sits between a forward and its backward.
                # The CPU generator is consumed in the reference's order (slice by slice, ascending instance id).
                plans = [losses.plan_instances(instance_labels[i * mb:min(B, (i + 1) * mb)],
                                               self.instance_loss.ignore_value) for i in range(n_slices)]
        total = None
        for i in range(n_slices):
            sl = slice(i * mb, min(B, (i + 1) * mb))
            if self.reducer is not None:
                self.reducer.accumulate(i < n_slices - 1)      # the all-reduce rides on the last slice's backward
            out = self.model(images[sl])
            logits = out[0] if isinstance(out, tuple) else out
            if panoptic:
                # the instance term is a mean over instances per slice, averaged over the slices
                loss = self._panoptic_backward(logits, labels[sl], instance_labels[sl], count, 1.0 / n_slices,
                                               plans[i] if plans is not None else None)
            else:
                loss = self._ce_backward(logits, labels[sl], count)
            total = loss if total is None else total + loss
        return total.detach()

    def _graph_step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        if self._graph is None:
            gx = torch.empty(images.shape, dtype=images.dtype, device=self.device)
            gy = torch.empty(labels.shape, dtype=labels.dtype, device=self.device)
            gx.copy_(images, non_blocking=True)
            gy.copy_(labels, non_blocking=True)
            self.optimizer.zero_grad(set_to_none=True)
            torch.cuda.synchronize(self.device)
            torch.cuda.empty_cache()             # the eager steps' activation cache goes back before the graph pool grows
            graph = torch.cuda.CUDAGraph()
            before = ops.LAUNCHES["count"]
            with torch.cuda.graph(graph):
                self.optimizer.zero_grad(set_to_none=True)
                loss = self.forward_backward(gx, gy)
                if self.reducer is not None:
                    self.reducer.finish()
                self.optimizer.step()
            self.graph_launches = ops.LAUNCHES["count"] - before
            self._graph = (graph, gx, gy, loss)
        graph, gx, gy, loss = self._graph
        if tuple(images.shape) != tuple(gx.shape) or tuple(labels.shape) != tuple(gy.shape):
            raise RuntimeError("Trainer(cuda_graph=True): the batch shape changed after the step was captured")
        gx.copy_(images, non_blocking=True)
        gy.copy_(labels, non_blocking=True)
        graph.replay()
        return loss.clone()

    def step(self, images: torch.Tensor, labels: torch.Tensor, instance_labels: Optional[torch.Tensor] = None,
             micro_batch: Optional[int] = None) -> torch.Tensor:
        """One training step; ``images``/``labels`` may live in (pinned) host memory.  Returns this rank's share of
        the loss (device tensor; mean over ranks = the global loss)."""
        if (self.cuda_graph and instance_labels is None and not (micro_batch and micro_batch < images.shape[0])
                and self.model.training):
            if self._eager_steps >= self.graph_warmup:
                return self._graph_step(images, labels)
            self._eager_steps += 1
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.forward_backward(images, labels, instance_labels, micro_batch)
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        return loss
