"""Generalised mode of the kernel sweep (BASELINE.json configs[4]).

The north star describes a Mask2Former-style block: per-query mask logits against pixel features
(``einsum('bqc,bnc->bqn')``), ``sigmoid > 0.5`` binarised into an additive attention bias, then masked multi-head
attention of Q queries over the N = H*W pixel tokens.  The reference module has none of this (its bias is
``torch.randint``, ade_semantic.py:177-181), so nothing here is on the drop-in path: it exists for
the sweep, and its oracle (oracle/query_attention_oracle.py) is builder-written -- parity unpinned by reference.

    bits, bits_t, row_count = query_mask_bits(qe, feat)                 # K13, tcgen05, logits never stored
    out = query_masked_attention(q, k, v, bits, bits_t, heads)          # K3 / K5 kernels, bias applied in registers

CUDA bf16 tensors only; no fallback.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import ops
from ._lib import MU_BF16, check

_L = ops._L
HEAD_DIM = 64            # the tcgen05 kernels' head width; 32-wide heads are zero-padded to it


def _bf16_cuda(*ts: Tensor) -> None:
    for t in ts:
        if not t.is_cuda or t.dtype != torch.bfloat16:
            raise RuntimeError("maskunet query-attention ops take CUDA bfloat16 tensors (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("maskunet ops need contiguous tensors")


@torch.library.custom_op("maskunet::query_mask_bits", mutates_args=(), device_types="cuda")
def query_mask_bits_op(qe: Tensor, feat: Tensor, want_logits: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _bf16_cuda(qe, feat)
    B, Q, C = qe.shape
    N = feat.shape[1]
    assert feat.shape == (B, N, C)
    NKP, QP = ops.nkp_of(N), ops.nkp_of(Q)
    dev = qe.device
    bits = torch.empty((B, Q, NKP // 32), dtype=torch.int32, device=dev)
    bits_t = torch.empty((B, NKP, QP // 32), dtype=torch.int32, device=dev)
    row_count = torch.empty((B, Q), dtype=torch.int32, device=dev)
    logits = torch.empty((B, Q, N) if want_logits else (0,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), ops._timed("mu_query_mask_bits", (B, Q, N, C)):
        ops._count(2)
        check(_L.mu_query_mask_bits(ops._p(qe), ops._p(feat), B, Q, N, C, ops._p(bits), ops._p(bits_t),
                                    ops._p(row_count), ops._optp(logits if want_logits else None), MU_BF16,
                                    ops._stream(qe)), "mu_query_mask_bits")
    return bits, bits_t, row_count, logits


@query_mask_bits_op.register_fake
def _(qe, feat, want_logits):
    B, Q, C = qe.shape
    N = feat.shape[1]
    i32 = dict(dtype=torch.int32)
    return (qe.new_empty((B, Q, ops.nkp_of(N) // 32), **i32), qe.new_empty((B, ops.nkp_of(N), ops.nkp_of(Q) // 32), **i32),
            qe.new_empty((B, Q), **i32), qe.new_empty((B, Q, N) if want_logits else (0,), dtype=torch.float32))


def query_mask_bits(qe: Tensor, feat: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """qe [B, Q, 256], feat [B, N, 256] bf16 -> (bits int32 [B, Q, NKP/32], bits_t int32 [B, NKP, QP/32],
    row_count int32 [B, Q]).  bit n%32 of bits[b, q, n/32] <=> sigmoid(<qe[b,q], feat[b,n]>) > 0.5; a query that kept
    nothing attends everything (row_count keeps the raw count)."""
    return query_mask_bits_op(qe.detach(), feat.detach(), False)[:3]


@torch.library.custom_op("maskunet::query_attn_fwd", mutates_args=(), device_types="cuda")
def query_attn_fwd(qh: Tensor, kh: Tensor, vh: Tensor, bits: Tensor, heads: int, n_keys: int, scale: float
                   ) -> Tuple[Tensor, Tensor]:
    """qh [BH, Q, 64], kh / vh [BH, NKP, 64] (rows >= n_keys zero), bits [BH / heads, Q, NKP/32] -> (o, lse)."""
    _bf16_cuda(qh, kh, vh)
    BH, Q, D = qh.shape
    NKP = kh.shape[1]
    o = torch.empty_like(qh)
    lse = torch.empty((BH, Q), dtype=torch.float32, device=qh.device)
    with torch.cuda.device(qh.device), ops._timed("mu_query_attn_fwd", (BH, Q, n_keys, D)):
        ops._count(1)
        check(_L.mu_query_attn_fwd(ops._p(qh), ops._p(kh), ops._p(vh), ops._p(bits), ops._p(o), ops._p(lse), BH, heads,
                                   Q, n_keys, NKP, D, scale, MU_BF16, ops._stream(qh)), "mu_query_attn_fwd")
    return o, lse


@query_attn_fwd.register_fake
def _(qh, kh, vh, bits, heads, n_keys, scale):
    return torch.empty_like(qh), qh.new_empty(qh.shape[:2], dtype=torch.float32)


@torch.library.custom_op("maskunet::query_attn_bwd", mutates_args=(), device_types="cuda")
def query_attn_bwd(qh: Tensor, kh: Tensor, vh: Tensor, bits_t: Tensor, d_o: Tensor, lse: Tensor, delta: Tensor,
                   heads: int, n_keys: int, scale: float) -> Tuple[Tensor, Tensor, Tensor]:
    """-> dq [BH, Q, 64], dk, dv [BH, NKP, 64] (rows >= n_keys zero)."""
    _bf16_cuda(qh, kh, vh, d_o)
    BH, Q, D = qh.shape
    NKP = kh.shape[1]
    dq = torch.empty_like(qh)
    dk = torch.zeros_like(kh)
    dv = torch.zeros_like(vh)
    with torch.cuda.device(qh.device):
        ws_bytes = int(_L.mu_query_attn_bwd_workspace_bytes(BH, Q, D))
        ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=qh.device)
        # dk / dv are [BH, n_keys, 64] for the kernel: hand it the first n_keys rows of every (sample, head) when the
        # key count is a multiple of the tile, otherwise a compact buffer copied back
        compact = n_keys != NKP
        dkc = torch.empty((BH, n_keys, D), dtype=qh.dtype, device=qh.device) if compact else dk
        dvc = torch.empty((BH, n_keys, D), dtype=qh.dtype, device=qh.device) if compact else dv
        with ops._timed("mu_query_attn_bwd", (BH, Q, n_keys, D)):
            ops._count(3)
            check(_L.mu_query_attn_bwd(ops._p(qh), ops._p(kh), ops._p(vh), ops._p(bits_t), ops._p(d_o), ops._p(lse),
                                       ops._p(delta), ops._p(dq), ops._p(dkc), ops._p(dvc), ops._p(ws), ws_bytes, BH,
                                       heads, Q, n_keys, NKP, D, scale, MU_BF16, ops._stream(qh)), "mu_query_attn_bwd")
        if compact:
            dk[:, :n_keys] = dkc
            dv[:, :n_keys] = dvc
    return dq, dk, dv


@query_attn_bwd.register_fake
def _(qh, kh, vh, bits_t, d_o, lse, delta, heads, n_keys, scale):
    return torch.empty_like(qh), torch.empty_like(kh), torch.empty_like(vh)


class _QueryAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qh, kh, vh, bits, bits_t, heads, n_keys, scale):
        o, lse = query_attn_fwd(qh, kh, vh, bits, heads, n_keys, scale)
        ctx.save_for_backward(qh, kh, vh, bits_t, o, lse)
        ctx.meta = (heads, n_keys, scale)
        return o

    @staticmethod
    def backward(ctx, d_o):
        qh, kh, vh, bits_t, o, lse = ctx.saved_tensors
        heads, n_keys, scale = ctx.meta
        d_o = d_o.contiguous()
        delta = (d_o.float() * o.float()).sum(-1)
        dq, dk, dv = query_attn_bwd(qh, kh, vh, bits_t, d_o, lse, delta, heads, n_keys, scale)
        return dq, dk, dv, None, None, None, None, None


def _to_heads(x: Tensor, heads: int, rows: int) -> Tensor:
    """[B, R, C] -> [B * heads, rows, 64]: head-major, rows and head width zero-padded."""
    B, R, C = x.shape
    d = C // heads
    x = x.view(B, R, heads, d).permute(0, 2, 1, 3)
    return F.pad(x, (0, HEAD_DIM - d, 0, rows - R)).reshape(B * heads, rows, HEAD_DIM).contiguous()


def query_masked_attention(q: Tensor, k: Tensor, v: Tensor, bits: Tensor, bits_t: Tensor, heads: int) -> Tensor:
    """out[b, q, h*d:(h+1)*d] = softmax_n(q_h k_h^T / sqrt(d) + bias[b, q, n]) v_h with bias = 0 where the bit of
    (q, n) is set and -inf elsewhere.  q [B, Q, C], k / v [B, N, C] bf16, C = heads * d, d in {32, 64}."""
    B, Q, C = q.shape
    N = k.shape[1]
    d = C // heads
    if d * heads != C or d not in (32, 64):
        raise ValueError("query_masked_attention: head width must be 32 or 64")
    NKP = ops.nkp_of(N)
    qh, kh, vh = _to_heads(q, heads, Q), _to_heads(k, heads, NKP), _to_heads(v, heads, NKP)
    o = _QueryAttn.apply(qh, kh, vh, bits, bits_t, heads, N, float(d) ** -0.5)
    return o[..., :d].reshape(B, heads, Q, d).permute(0, 2, 1, 3).reshape(B, Q, C)
