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
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import ops
from .ddp import GradReducer


class Trainer:
    def __init__(self, model: torch.nn.Module, lr: float = 5e-5, weight_decay: float = 1e-1,
                 ignore_index: int = -100, data_parallel: bool = False, bucket_bytes: int = 25 * 1024 * 1024,
                 instance_loss: Optional[torch.nn.Module] = None, loss_weights=(0.9, 0.1)):
        self.model = model
        self.device = next(model.parameters()).device
        self.ignore_index = ignore_index
        self.instance_loss = instance_loss          # maskunet_b200.InstanceContrastiveLoss or None
        self.loss_weights = loss_weights            # (semantic, instance), coco_panoptic.py:552
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, fused=self.device.type == "cuda")
        self.reducer: Optional[GradReducer] = None
        if data_parallel:
            self.reducer = GradReducer(params, bucket_bytes=bucket_bytes)
            self.reducer.broadcast_parameters(model)

    def _step_with_instance_term(self, logits, labels, inst):
        from . import losses
        w_sem, w_inst = self.loss_weights
        crit = self.instance_loss
        if inst.device != self.device:
            inst = inst.to(self.device, non_blocking=True)
        padded = getattr(self.model, "_padded_logits", None)
        fused = (logits.is_cuda and isinstance(crit, losses.InstanceContrastiveLoss) and padded is not None
                 and padded.requires_grad and padded.data_ptr() == logits.data_ptr()
                 and padded.is_contiguous(memory_format=torch.channels_last) and logits.shape[1] <= 256)
        if not fused:
            loss = (w_sem * F.cross_entropy(logits.float(), labels, ignore_index=self.ignore_index)
                    + w_inst * crit(logits, inst))
            loss.backward()
            return loss
        loss, dpad = self.fused_panoptic_loss(padded, logits.shape[1], labels, inst)
        padded.backward(dpad)
        return loss

    def fused_panoptic_loss(self, padded, c_out, labels, inst):
        """(w_sem * CE + w_inst * InstanceContrastiveLoss, its gradient w.r.t. the class-padded logit buffer).  Both
        gradients land in ONE buffer: the fused cross-entropy writes d(CE)/d(logits) into it, the triplet kernel adds
        its three logit columns per instance; backward then starts from the sum."""
        from . import losses
        w_sem, w_inst = self.loss_weights
        crit = self.instance_loss
        with torch.no_grad():
            padded = padded.detach()
            loss, dpad = ops.cross_entropy_fused(padded, labels, self.ignore_index, c_out)
            dpad.mul_(w_sem)
            loss = w_sem * loss.squeeze(0)
            order, meta, K = losses.plan_instances(inst, crit.ignore_value)
            if K:
                view = padded[:, :c_out]
                l_inst, sel, dist = losses.instance_triplet(view, order, meta, float(crit.margin))
                losses.accumulate_grad(view, sel, dist, float(crit.margin), dpad[:, :c_out], scale=w_inst)
                loss = loss + w_inst * l_inst.squeeze(0)
        return loss, dpad

    def step(self, images: torch.Tensor, labels: torch.Tensor,
             instance_labels: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step; ``images``/``labels`` may live in (pinned) host memory.  Returns the loss (device)."""
        if images.device != self.device:
            images = images.to(self.device, non_blocking=True)
        if labels.device != self.device:
            labels = labels.to(self.device, non_blocking=True)
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(images)
        logits = out[0] if isinstance(out, tuple) else out
        if self.instance_loss is not None and instance_labels is not None:
            loss = self._step_with_instance_term(logits, labels, instance_labels)
            if self.reducer is not None:
                self.reducer.finish()
            self.optimizer.step()
            return loss.detach()
        fused_ok = logits.is_cuda and logits.shape[1] <= 256 and logits.dtype in (torch.float32, torch.bfloat16)
        padded = getattr(self.model, "_padded_logits", None)
        if (fused_ok and padded is not None and padded.requires_grad and padded.data_ptr() == logits.data_ptr()
                and padded.is_contiguous(memory_format=torch.channels_last)):
            # logits are the first c_out channels of the head's class-padded buffer: run the fused loss on the
            # buffer itself and start backward there (no narrow / pad passes, no loss node)
            with torch.no_grad():
                loss, dpad = ops.cross_entropy_fused(padded.detach(), labels, self.ignore_index, logits.shape[1])
            loss = loss.squeeze(0)
            padded.backward(dpad)
        elif fused_ok and logits.is_contiguous(memory_format=torch.channels_last):
            # fused CrossEntropyLoss(mean) forward + gradient on channels-last logits (one read, one write), and
            # backward starts from d(loss)/d(logits) directly: the loss node and its `dlogits * dloss` pass are skipped
            with torch.no_grad():
                loss, dlogits = ops.cross_entropy_fused(logits.detach(), labels, self.ignore_index)
            loss = loss.squeeze(0)
            logits.backward(dlogits)
        else:
            loss = F.cross_entropy(logits.float(), labels, ignore_index=self.ignore_index)
            loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.optimizer.step()
        return loss.detach()
