"""CPU restatement of the reference Mask Attention Module.  TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/code/ade20k/ade_semantic.py:152-190`` (class
``Mask2FormerAttention``; eight byte-equivalent copies elsewhere, SURVEY.md
section 2).  Written as explicit tensor algebra with a hand-derived backward so
that it is an independent statement of the maths, not a call into the same
autograd graph the reference builds.

Pinned against the reference's own class by ``tests/golden/make_golden.py``
(goldens in ``tests/golden/attn_*.npz``) and, when ``/root/reference`` is
present, live in ``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

NEG_INF = float("-inf")


# --------------------------------------------------------------------------- mask
def draw_mask_bits(batch: int, height: int, width: int, device="cpu") -> torch.Tensor:
    """The one RNG-consuming call of the module: ade_semantic.py:178.

    Same call, same argument order, so the torch global generator is consumed
    exactly as the reference consumes it.
    """
    return torch.randint(0, 2, (batch, height, width), device=device)


def binarize_mask(bits: torch.Tensor) -> torch.Tensor:
    """ade_semantic.py:179-180: ``binary_mask > 0.5`` on the flattened int64 draw.

    Returns keep[B, N] (bool): True -> additive bias 0.0, False -> -inf.
    """
    flat = bits.reshape(bits.shape[0], -1)
    return flat > 0.5


def additive_bias(keep: torch.Tensor) -> torch.Tensor:
    """ade_semantic.py:180: where(keep, 0.0, -inf) as fp32 [B, N]."""
    zero = torch.zeros((), dtype=torch.float32, device=keep.device)
    ninf = torch.full((), NEG_INF, dtype=torch.float32, device=keep.device)
    return torch.where(keep, zero, ninf)


def expand_bias(bias: torch.Tensor, n_query: int) -> torch.Tensor:
    """ade_semantic.py:181: unsqueeze(1).expand(-1, N, -1) -> [B, N, N], stride (N, 0, 1)."""
    return bias.unsqueeze(1).expand(-1, n_query, -1)


def keep_from_module_mask(mask: torch.Tensor) -> torch.Tensor:
    """Recover keep[B, N] from a cached ``self.mask`` ([B, N, N] expanded 0/-inf)."""
    return mask[:, 0, :] == 0


# --------------------------------------------------------------------------- forward
def attention_forward(x: torch.Tensor, params: Dict[str, torch.Tensor], keep: torch.Tensor,
                      eps: float = 1e-5, dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Forward of the module, ade_semantic.py:163-190.

    x      [B, C, H, W] contiguous
    params query/key/value.{weight,bias}, norm.{weight,bias}  (state_dict keys)
    keep   [B, N] bool, N = H*W (per-key, shared by every query of the sample)

    Returns dict with ``y`` ([B, N, C] row-major; the module returns these bytes
    reinterpreted as [B, C, H, W], line :190) plus everything backward needs.
    """
    B, C, H, W = x.shape
    N = H * W
    xt = x.to(dtype).reshape(B, C, N).transpose(1, 2)            # :168  tokens [B, N, C]
    wq, bq = params["query.weight"].to(dtype), params["query.bias"].to(dtype)
    wk, bk = params["key.weight"].to(dtype), params["key.bias"].to(dtype)
    wv, bv = params["value.weight"].to(dtype), params["value.bias"].to(dtype)
    gamma, beta = params["norm.weight"].to(dtype), params["norm.bias"].to(dtype)

    q = xt @ wq.t() + bq                                          # :170
    k = xt @ wk.t() + bk                                          # :171
    v = xt @ wv.t() + bv                                          # :172
    s = (q @ k.transpose(1, 2)) / (C ** 0.5)                      # :174-175 (true division)
    bias = torch.where(keep, torch.zeros((), dtype=dtype), torch.full((), NEG_INF, dtype=dtype))
    s = s + bias.unsqueeze(1)                                     # :183 (broadcast over queries)
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    l = e.sum(dim=-1, keepdim=True)
    p = e / l                                                     # :185
    o = p @ v                                                     # :186
    z = o + xt                                                    # :187
    mu = z.mean(dim=-1, keepdim=True)
    var = ((z - mu) ** 2).mean(dim=-1, keepdim=True)              # biased, as nn.LayerNorm
    rstd = torch.rsqrt(var + eps)
    zhat = (z - mu) * rstd
    y = zhat * gamma + beta                                       # :188
    lse = (m + torch.log(l)).squeeze(-1)
    return dict(y=y, xt=xt, q=q, k=k, v=v, p=p, o=o, zhat=zhat, rstd=rstd, lse=lse,
                wq=wq, wk=wk, wv=wv, gamma=gamma)


def module_output(y: torch.Tensor, C: int, H: int, W: int) -> torch.Tensor:
    """ade_semantic.py:190: ``.view(B, C, H, W)`` of the [B, N, C] buffer (no permute back)."""
    return y.contiguous().view(y.shape[0], C, H, W)


# --------------------------------------------------------------------------- backward
def attention_backward(saved: Dict[str, torch.Tensor], dy: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Hand-derived gradients of ``attention_forward`` wrt x and the 8 parameters.

    dy is the gradient wrt y in its [B, N, C] layout (i.e. the incoming
    [B, C, H, W] gradient reinterpreted, mirroring :190).  What autograd does
    implicitly for the reference at ade_semantic.py:400.
    """
    xt, q, k, v, p = saved["xt"], saved["q"], saved["k"], saved["v"], saved["p"]
    zhat, rstd, gamma = saved["zhat"], saved["rstd"], saved["gamma"]
    B, N, C = xt.shape
    dy = dy.to(xt.dtype).reshape(B, N, C)

    dgamma = (dy * zhat).sum(dim=(0, 1))
    dbeta = dy.sum(dim=(0, 1))
    g = dy * gamma
    dz = rstd * (g - g.mean(dim=-1, keepdim=True) - zhat * (g * zhat).mean(dim=-1, keepdim=True))

    do = dz                                   # through  z = o + xt
    dxt = dz.clone()
    dv = p.transpose(1, 2) @ do
    dp = do @ v.transpose(1, 2)
    delta = (dp * p).sum(dim=-1, keepdim=True)
    ds = p * (dp - delta) / (C ** 0.5)        # masked keys: p == 0 exactly -> ds == 0
    dq = ds @ k
    dk = ds.transpose(1, 2) @ q

    out = {}
    for name, d, w in (("query", dq, saved["wq"]), ("key", dk, saved["wk"]), ("value", dv, saved["wv"])):
        out[f"{name}.weight"] = torch.einsum("bno,bni->oi", d, xt)
        out[f"{name}.bias"] = d.sum(dim=(0, 1))
        dxt = dxt + d @ w
    out["norm.weight"] = dgamma
    out["norm.bias"] = dbeta
    out["x_tokens"] = dxt                      # [B, N, C]
    out["x"] = dxt.transpose(1, 2).contiguous()  # [B, C, N] == NCHW flattened
    return out


def useful_flops(n_tokens: int, n_keep: int, channels: int) -> Dict[str, int]:
    """Algorithmic work per image per site, SURVEY.md section 8(d)."""
    return dict(attn_fwd=4 * n_tokens * n_keep * channels,
                attn_bwd=8 * n_tokens * n_keep * channels,
                proj_fwd=6 * n_tokens * channels * channels,
                dense_fwd=4 * n_tokens * n_tokens * channels)


def sdpa_reference(q, k, v, keep, channels: Optional[int] = None):
    """Kernel-level oracle: softmax(q k^T / sqrt(C) + bias) v on given q, k, v."""
    C = channels or q.shape[-1]
    s = (q @ k.transpose(1, 2)) / math.sqrt(C)
    s = s.masked_fill(~keep.unsqueeze(1), NEG_INF)
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    l = e.sum(dim=-1, keepdim=True)
    return (e / l) @ v, (m + torch.log(l)).squeeze(-1)
