"""CPU oracle of the generalised mode: per-query mask logits, sigmoid > 0.5 binarisation,
masked multi-head attention of Q queries over N pixel tokens.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED BY REFERENCE: the reference has no such stage (its attention bias is torch.randint,
/root/reference/code/ade20k/ade_semantic.py:177-181; SURVEY.md section 0 lists the north-star vocabulary that the
reference does not contain).  This file is builder-written plain PyTorch of the published Mask2Former recipe
(masked attention with ``attn_mask = sigmoid(mask_logits) < 0.5`` and "a query whose mask is empty attends
everywhere"); it pins the CUDA kernels to that recipe, not to the reference.
"""
from __future__ import annotations

import torch


def mask_logits(qe: torch.Tensor, feat: torch.Tensor) -> torch.Tensor:
    """einsum('bqc,bnc->bqn') in fp32 (inputs are bf16-representable in the tests)."""
    return torch.einsum("bqc,bnc->bqn", qe.float(), feat.float())


def keep_from_logits(logits: torch.Tensor):
    """(keep bool [B, Q, N] after the empty-row rule, raw kept count per row)."""
    keep = torch.sigmoid(logits.float()) > 0.5
    count = keep.sum(-1)
    keep = keep | (count == 0).unsqueeze(-1)
    return keep, count


def pack_bits(keep: torch.Tensor, words: int) -> torch.Tensor:
    """bool [..., N] -> int32 [..., words], bit n % 32 of word n // 32 (little end first), zero padded."""
    *lead, N = keep.shape
    k = torch.zeros(*lead, words * 32, dtype=torch.int64)
    k[..., :N] = keep.to(torch.int64)
    w = (k.view(*lead, words, 32) << torch.arange(32, dtype=torch.int64)).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)          # two's complement view of the uint32 word


def unpack_bits(bits: torch.Tensor, n: int) -> torch.Tensor:
    """int32 [..., words] -> bool [..., n]."""
    w = bits.to(torch.int64) & 0xFFFFFFFF
    b = (w.unsqueeze(-1) >> torch.arange(32, dtype=torch.int64)) & 1
    return b.reshape(*bits.shape[:-1], -1)[..., :n].bool()


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, keep: torch.Tensor, heads: int) -> torch.Tensor:
    """q [B, Q, C], k / v [B, N, C], keep bool [B, Q, N] shared by the heads -> [B, Q, C]; fp32 throughout."""
    B, Q, C = q.shape
    N = k.shape[1]
    d = C // heads
    qh = q.float().view(B, Q, heads, d).transpose(1, 2)
    kh = k.float().view(B, N, heads, d).transpose(1, 2)
    vh = v.float().view(B, N, heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / d ** 0.5
    s = s.masked_fill(~keep.unsqueeze(1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Q, C)
