"""The reference's timed loop body as a reusable step.

/root/reference/code/ade20k/ade_semantic.py:394-401:
    inputs, labels = inputs.to(device), labels.to(device)
    optimizer.zero_grad(); outputs = model(inputs); loss = criterion(outputs, labels)
    loss.backward(); optimizer.step()
with ``criterion = nn.CrossEntropyLoss()`` and ``optim.AdamW(lr=5e-5, weight_decay=1e-1)`` (:377-379).
Under data parallelism the gradient all-reduce (ddp.GradReducer) sits between backward and the optimiser.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import ops
from .ddp import GradReducer


class Trainer:
    def __init__(self, model: torch.nn.Module, lr: float = 5e-5, weight_decay: float = 1e-1,
                 ignore_index: int = -100, data_parallel: bool = False, bucket_bytes: int = 25 * 1024 * 1024):
        self.model = model
        self.device = next(model.parameters()).device
        self.ignore_index = ignore_index
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, fused=self.device.type == "cuda")
        self.reducer: Optional[GradReducer] = None
        if data_parallel:
            self.reducer = GradReducer(params, bucket_bytes=bucket_bytes)
            self.reducer.broadcast_parameters(model)

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """One training step; ``images``/``labels`` may live in (pinned) host memory.  Returns the loss (device)."""
        if images.device != self.device:
            images = images.to(self.device, non_blocking=True)
        if labels.device != self.device:
            labels = labels.to(self.device, non_blocking=True)
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(images)
        logits = out[0] if isinstance(out, tuple) else out
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
