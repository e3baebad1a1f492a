"""torch.library operators (namespace ``maskunet::``) over the C ABI.

Each op allocates its outputs with the PyTorch caching allocator, passes raw
device pointers and the current CUDA stream to libmaskunet_b200.so and converts
a non-zero status into RuntimeError(mu_last_error()).  CUDA tensors only.

Layouts (B batch, C channels, N = H*W tokens, NKP = roundup(N, 128)):
  x   [B, C, N]   channel-major (NCHW flattened)            -> CUDA-core projection / LN kernels
      [B, N, C]   token-major (channels-last, bf16 only)    -> tcgen05 projection kernels
  q, o, y, dz, dq, dk, dv   [B, N, C]
  kc, vc                    [B, NKP, C]  K / V rows of the kept keys only
  w_qkv [3C, C], b_qkv [3C]              cat(query, key, value) parameters, fp32
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import MU_BF16, MU_F32, check

_L = _lib.load()

# ---- instrumentation used by bench.py (not by the product path) ---------------------------------
LAUNCHES = {"count": 0}          # kernels of OURS launched so far (each C-ABI call adds what it enqueues)
KERNEL_TIMING = {"enabled": False, "events": {}}   # name -> [(start_event, end_event, meta), ...]


def _count(n: int) -> None:
    LAUNCHES["count"] += n


class _timed:
    """Bracket one C-ABI call with CUDA events on the launching stream when timing is enabled."""

    def __init__(self, name, meta=None):
        self.name, self.meta = name, meta

    def __enter__(self):
        if KERNEL_TIMING["enabled"]:
            self.t0 = torch.cuda.Event(enable_timing=True)
            self.t1 = torch.cuda.Event(enable_timing=True)
            self.t0.record()
        return self

    def __exit__(self, *exc):
        if KERNEL_TIMING["enabled"]:
            self.t1.record()
            KERNEL_TIMING["events"].setdefault(self.name, []).append((self.t0, self.t1, self.meta))
        return False


def _p(t: Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _stream(t: Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _code(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return MU_F32
    if t.dtype == torch.bfloat16:
        return MU_BF16
    raise TypeError(f"maskunet ops take float32 or bfloat16 activations, got {t.dtype}")


def _cuda(*ts: Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("maskunet ops run on CUDA tensors only (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("maskunet ops need contiguous tensors")


def nkp_of(n: int) -> int:
    return (n + 127) // 128 * 128


def _bnc(x: Tensor, token_major: bool):
    """(B, C, N) of an activation in either layout."""
    if token_major:
        return x.shape[0], x.shape[2], x.shape[1]
    return x.shape[0], x.shape[1], x.shape[2]


# ------------------------------------------------------------------ K2
@torch.library.custom_op("maskunet::mask_binarize", mutates_args=(), device_types="cuda")
def mask_binarize(bits: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """bits int64 [B, N] -> keep_bits int32 [B, ceil(N/32)], n_keep int32 [B], keep_idx, keep_rank int32 [B, N]."""
    _cuda(bits)
    assert bits.dtype == torch.int64 and bits.dim() == 2
    B, N = bits.shape
    dev = bits.device
    keep_bits = torch.empty((B, (N + 31) // 32), dtype=torch.int32, device=dev)
    n_keep = torch.empty((B,), dtype=torch.int32, device=dev)
    keep_idx = torch.empty((B, N), dtype=torch.int32, device=dev)
    keep_rank = torch.empty((B, N), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _count(1)
        check(_L.mu_mask_binarize(_p(bits), B, N, _p(keep_bits), _p(n_keep), _p(keep_idx), _p(keep_rank),
                                  _stream(bits)), "mu_mask_binarize")
    return keep_bits, n_keep, keep_idx, keep_rank


@mask_binarize.register_fake
def _(bits):
    B, N = bits.shape
    i32 = dict(dtype=torch.int32, device=bits.device)
    return (torch.empty((B, (N + 31) // 32), **i32), torch.empty((B,), **i32),
            torch.empty((B, N), **i32), torch.empty((B, N), **i32))


# ------------------------------------------------------------------ K1
@torch.library.custom_op("maskunet::qkv_project", mutates_args=(), device_types="cuda")
def qkv_project(x: Tensor, w_qkv: Tensor, b_qkv: Tensor, keep_rank: Tensor, n_keep: Tensor, token_major: bool
                ) -> Tuple[Tensor, Tensor, Tensor]:
    _cuda(x, w_qkv, b_qkv, keep_rank, n_keep)
    B, C, N = _bnc(x, token_major)
    NKP = nkp_of(N)
    q = torch.empty((B, N, C), dtype=x.dtype, device=x.device)
    kc = torch.empty((B, NKP, C), dtype=x.dtype, device=x.device)
    vc = torch.empty((B, NKP, C), dtype=x.dtype, device=x.device)
    w_lp = w_qkv.to(torch.bfloat16) if token_major else w_qkv
    with torch.cuda.device(x.device):
        _count(2)
        check(_L.mu_qkv_project(_p(x), _p(w_qkv), _p(w_lp), _p(b_qkv), _p(keep_rank), _p(n_keep), _p(q), _p(kc),
                                _p(vc), B, C, N, NKP, _code(x), int(token_major), _stream(x)), "mu_qkv_project")
    return q, kc, vc


@qkv_project.register_fake
def _(x, w_qkv, b_qkv, keep_rank, n_keep, token_major):
    B, C, N = _bnc(x, token_major)
    return x.new_empty((B, N, C)), x.new_empty((B, nkp_of(N), C)), x.new_empty((B, nkp_of(N), C))


# ------------------------------------------------------------------ K3
def _attn_fwd_impl(fn, name, q, kc, vc, n_keep):
    _cuda(q, kc, vc, n_keep)
    B, N, C = q.shape
    NKP = kc.shape[1]
    o = torch.empty_like(q)
    lse = torch.empty((B, N), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device), _timed(name, (B, N, C)):
        _count(1)
        check(fn(_p(q), _p(kc), _p(vc), _p(n_keep), _p(o), _p(lse), B, N, NKP, C, _code(q), _stream(q)), name)
    return o, lse


@torch.library.custom_op("maskunet::attn_fwd", mutates_args=(), device_types="cuda")
def attn_fwd(q: Tensor, kc: Tensor, vc: Tensor, n_keep: Tensor) -> Tuple[Tensor, Tensor]:
    """O = softmax(Q Kc^T / sqrt(C)) Vc over kept keys; bf16 -> tcgen05 kernel, fp32 -> CUDA cores."""
    return _attn_fwd_impl(_L.mu_attn_fwd, "mu_attn_fwd", q, kc, vc, n_keep)


@attn_fwd.register_fake
def _(q, kc, vc, n_keep):
    return torch.empty_like(q), q.new_empty(q.shape[:2], dtype=torch.float32)


def attn_fwd_cudacore(q, kc, vc, n_keep):
    """Diagnostic: the CUDA-core kernel for either dtype (cross-check of the tcgen05 kernel)."""
    return _attn_fwd_impl(_L.mu_attn_fwd_cudacore, "mu_attn_fwd_cudacore", q, kc, vc, n_keep)


# ------------------------------------------------------------------ K5
def _attn_bwd_impl(tensor_core, q, kc, vc, n_keep, keep_idx, d_o, lse, delta):
    _cuda(q, kc, vc, n_keep, keep_idx, d_o, lse, delta)
    B, N, C = q.shape
    NKP = kc.shape[1]
    dq = torch.empty_like(q)
    dk = torch.empty_like(q)     # token space; the entry point clears dk / dv before scattering rows
    dv = torch.empty_like(q)
    name = "mu_attn_bwd" if tensor_core else "mu_attn_bwd_cudacore"
    with torch.cuda.device(q.device):
        if tensor_core:
            ws_bytes = int(_L.mu_attn_bwd_workspace_bytes(B, N, C, _code(q)))
            ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=q.device)
            with _timed(name, (B, N, C)):
                _count(1 if C == 64 else 2)      # d = 64 adds bf16 partial tiles straight into dq: no convert kernel
                check(_L.mu_attn_bwd(_p(q), _p(kc), _p(vc), _p(n_keep), _p(keep_idx), _p(d_o), _p(lse), _p(delta),
                                     _p(dq), _p(dk), _p(dv), _p(ws), ws_bytes, B, N, NKP, C, _code(q), _stream(q)),
                      name)
        else:
            with _timed(name, (B, N, C)):
                _count(2)
                check(_L.mu_attn_bwd_cudacore(_p(q), _p(kc), _p(vc), _p(n_keep), _p(keep_idx), _p(d_o), _p(lse),
                                              _p(delta), _p(dq), _p(dk), _p(dv), B, N, NKP, C, _code(q), _stream(q)),
                      name)
    return dq, dk, dv


@torch.library.custom_op("maskunet::attn_bwd", mutates_args=(), device_types="cuda")
def attn_bwd(q: Tensor, kc: Tensor, vc: Tensor, n_keep: Tensor, keep_idx: Tensor, d_o: Tensor, lse: Tensor,
             delta: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """dq, dk, dv, all [B, N, C] in token space (rows of masked keys are zero)."""
    return _attn_bwd_impl(True, q, kc, vc, n_keep, keep_idx, d_o, lse, delta)


@attn_bwd.register_fake
def _(q, kc, vc, n_keep, keep_idx, d_o, lse, delta):
    return torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)


def attn_bwd_cudacore(q, kc, vc, n_keep, keep_idx, d_o, lse, delta):
    return _attn_bwd_impl(False, q, kc, vc, n_keep, keep_idx, d_o, lse, delta)


# ------------------------------------------------------------------ K3 epilogue / K4
@torch.library.custom_op("maskunet::residual_ln_fwd", mutates_args=(), device_types="cuda")
def residual_ln_fwd(o: Tensor, x: Tensor, gamma: Tensor, beta: Tensor, eps: float, token_major: bool,
                    view_out: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    _cuda(o, x, gamma, beta)
    B, N, C = o.shape
    y = torch.empty_like(o)
    mean = torch.empty((B, N), dtype=torch.float32, device=o.device)
    rstd = torch.empty((B, N), dtype=torch.float32, device=o.device)
    with torch.cuda.device(o.device):
        _count(1)
        check(_L.mu_residual_ln_fwd(_p(o), _p(x), _p(gamma), _p(beta), eps, _p(y), _p(mean), _p(rstd),
                                    B, C, N, _code(o), 2 if view_out else int(token_major), _stream(o)),
              "mu_residual_ln_fwd")
    return y, mean, rstd


@residual_ln_fwd.register_fake
def _(o, x, gamma, beta, eps, token_major, view_out=False):
    s = o.new_empty(o.shape[:2], dtype=torch.float32)
    return torch.empty_like(o), s, torch.empty_like(s)


@torch.library.custom_op("maskunet::residual_ln_bwd", mutates_args=(), device_types="cuda")
def residual_ln_bwd(dy: Tensor, o: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, gamma: Tensor, token_major: bool,
                    view_out: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    _cuda(dy, o, x, mean, rstd, gamma)
    B, N, C = o.shape
    dz = torch.empty_like(o)
    delta = torch.empty((B, N), dtype=torch.float32, device=o.device)
    dgamma = torch.zeros((C,), dtype=torch.float32, device=o.device)
    dbeta = torch.zeros((C,), dtype=torch.float32, device=o.device)
    with torch.cuda.device(o.device):
        _count(1)
        check(_L.mu_residual_ln_bwd(_p(dy), _p(o), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dz), _p(delta),
                                    _p(dgamma), _p(dbeta), B, C, N, _code(o), 2 if view_out else int(token_major),
                                    _stream(o)),
              "mu_residual_ln_bwd")
    return dz, delta, dgamma, dbeta


@residual_ln_bwd.register_fake
def _(dy, o, x, mean, rstd, gamma, token_major, view_out=False):
    C = o.shape[-1]
    return (torch.empty_like(o), o.new_empty(o.shape[:2], dtype=torch.float32),
            o.new_empty((C,), dtype=torch.float32), o.new_empty((C,), dtype=torch.float32))


# ------------------------------------------------------------------ K6
@torch.library.custom_op("maskunet::qkv_project_bwd", mutates_args=(), device_types="cuda")
def qkv_project_bwd(x: Tensor, dz: Tensor, dq: Tensor, dk: Tensor, dv: Tensor, w_qkv: Tensor, token_major: bool
                    ) -> Tuple[Tensor, Tensor, Tensor]:
    _cuda(x, dz, dq, dk, dv, w_qkv)
    B, C, N = _bnc(x, token_major)
    dx = torch.empty_like(x)
    dw = torch.zeros((3 * C, C), dtype=torch.float32, device=x.device)
    db = torch.zeros((3 * C,), dtype=torch.float32, device=x.device)
    w_lp = w_qkv.to(torch.bfloat16) if token_major else w_qkv
    with torch.cuda.device(x.device):
        # P2, P3 (+ the bias-gradient kernel in deterministic mode: otherwise db is a column of the P3 GEMM)
        _count((3 if _L.mu_get_deterministic() else 2) if token_major else 2)
        check(_L.mu_qkv_project_bwd(_p(x), _p(dz), _p(dq), _p(dk), _p(dv), _p(w_qkv), _p(w_lp), _p(dx), _p(dw),
                                    _p(db), B, C, N, _code(x), int(token_major), _stream(x)), "mu_qkv_project_bwd")
    return dx, dw, db


@qkv_project_bwd.register_fake
def _(x, dz, dq, dk, dv, w_qkv, token_major):
    C = w_qkv.shape[1]
    return torch.empty_like(x), x.new_empty((3 * C, C), dtype=torch.float32), x.new_empty((3 * C,), dtype=torch.float32)


# ------------------------------------------------------------------ layout bridge
@torch.library.custom_op("maskunet::transpose", mutates_args=(), device_types="cuda")
def transpose(x: Tensor) -> Tensor:
    """[batch, rows, cols] -> [batch, cols, rows] (2- or 4-byte elements), coalesced on both sides."""
    _cuda(x)
    Bt, R, Cc = x.shape
    out = torch.empty((Bt, Cc, R), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _count(1)
        check(_L.mu_transpose(_p(x), _p(out), Bt, R, Cc, x.element_size(), _stream(x)), "mu_transpose")
    return out


@transpose.register_fake
def _(x):
    return x.new_empty((x.shape[0], x.shape[2], x.shape[1]))


transpose.register_autograd(lambda ctx, g: transpose(g.contiguous()))


# ------------------------------------------------------------------ K8: fused BatchNorm + activation (+ residual)
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def _rows(x: Tensor):
    """x is a 4-D channels-last tensor: its memory is [M = B*H*W, C] rows."""
    if not (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        raise RuntimeError("maskunet bn_act ops take 4-D channels-last tensors")
    return x.shape[0] * x.shape[2] * x.shape[3], x.shape[1]


def _optp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


@torch.library.custom_op("maskunet::bn_act_fwd", mutates_args=(), device_types="cuda")
def bn_act_fwd(x: Tensor, residual: Tensor | None, gamma: Tensor, beta: Tensor, eps: float, act: int
               ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Training-mode BN over (B, H, W) + activation (+ residual).  Returns y, mean, rstd, a, b.
    (Functional: the caller folds mean / rstd into the module's running statistics.)"""
    M, C = _rows(x)
    if residual is not None:
        _rows(residual)
    y = torch.empty_like(x)
    f32 = dict(dtype=torch.float32, device=x.device)
    mean, rstd, a, b = (torch.empty((C,), **f32) for _ in range(4))
    sums = torch.empty((2 * C,), **f32)
    with torch.cuda.device(x.device):
        _count(3)
        check(_L.mu_bn_act_fwd(_p(x), _optp(residual), _p(gamma), _p(beta), _optp(None), _optp(None),
                               0.0, eps, _p(y), _p(mean), _p(rstd), _p(a), _p(b), _p(sums), M, C, act, _code(x),
                               _stream(x)), "mu_bn_act_fwd")
    return y, mean, rstd, a, b


@bn_act_fwd.register_fake
def _(x, residual, gamma, beta, eps, act):
    C = x.shape[1]
    v = lambda: x.new_empty((C,), dtype=torch.float32)
    return torch.empty_like(x), v(), v(), v(), v()


@torch.library.custom_op("maskunet::bn_act_apply", mutates_args=(), device_types="cuda")
def bn_act_apply(x: Tensor, residual: Tensor | None, a: Tensor, b: Tensor, act: int) -> Tensor:
    """y = act(a * x + b [+ residual]) with per-channel a, b (inference-mode BN folded to an affine)."""
    M, C = _rows(x)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _count(1)
        check(_L.mu_bn_act_apply(_p(x), _optp(residual), _p(a), _p(b), _p(y), M, C, act, _code(x), _stream(x)),
              "mu_bn_act_apply")
    return y


@bn_act_apply.register_fake
def _(x, residual, a, b, act):
    return torch.empty_like(x)


@torch.library.custom_op("maskunet::bn_act_bwd", mutates_args=(), device_types="cuda")
def bn_act_bwd(dy: Tensor, x: Tensor, residual: Tensor | None, a: Tensor, b: Tensor, mean: Tensor, rstd: Tensor,
               act: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Returns dx, dresidual (empty tensor when there was no residual), dgamma, dbeta."""
    M, C = _rows(x)
    _rows(dy)
    dx = torch.empty_like(x)
    dr = torch.empty_like(x) if residual is not None else None
    sums = torch.empty((2 * C,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _count(2)
        check(_L.mu_bn_act_bwd(_p(dy), _p(x), _optp(residual), _p(a), _p(b), _p(mean), _p(rstd), _p(sums), _p(dx),
                               _optp(dr), M, C, act, _code(x), _stream(x)), "mu_bn_act_bwd")
    if dr is None:
        dr = x.new_empty((0,))
    return dx, dr, sums[C:].clone(), sums[:C].clone()


@bn_act_bwd.register_fake
def _(dy, x, residual, a, b, mean, rstd, act):
    C = x.shape[1]
    dr = torch.empty_like(x) if residual is not None else x.new_empty((0,))
    return torch.empty_like(x), dr, x.new_empty((C,), dtype=torch.float32), x.new_empty((C,), dtype=torch.float32)


def _bn_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)   # unused outputs get None, not dense zero gradients
    x, residual, gamma, beta, eps, act = inputs
    y, mean, rstd, a, b = output
    ctx.act = act
    ctx.has_res = residual is not None
    if residual is not None:
        ctx.save_for_backward(x, residual, a, b, mean, rstd)
    else:
        ctx.save_for_backward(x, a, b, mean, rstd)


def _bn_backward(ctx, dy, *unused):
    if ctx.has_res:
        x, residual, a, b, mean, rstd = ctx.saved_tensors
    else:
        (x, a, b, mean, rstd), residual = ctx.saved_tensors, None
    dy = dy.contiguous(memory_format=torch.channels_last)
    dx, dr, dgamma, dbeta = bn_act_bwd(dy, x, residual, a, b, mean, rstd, ctx.act)
    return dx, (dr if ctx.has_res else None), dgamma, dbeta, None, None


bn_act_fwd.register_autograd(_bn_backward, setup_context=_bn_setup)


@torch.library.custom_op("maskunet::bn_update_running", mutates_args=("running_mean", "running_var"),
                         device_types="cuda")
def bn_update_running(running_mean: Tensor, running_var: Tensor, mean: Tensor, rstd: Tensor, momentum: float,
                      eps: float, count: int) -> None:
    """nn.BatchNorm2d's running-statistics update (momentum, unbiased variance) from the batch mean / rstd that
    bn_act_fwd returned: one small kernel instead of six elementwise launches per layer."""
    _cuda(running_mean, running_var, mean, rstd)
    C = mean.numel()
    with torch.cuda.device(mean.device):
        _count(1)
        check(_L.mu_bn_update_running(_p(running_mean), _p(running_var), _p(mean), _p(rstd), momentum, eps, count, C,
                                      _stream(mean)), "mu_bn_update_running")


# ------------------------------------------------------------------ K8 with the statistics pass done by the producer
@torch.library.custom_op("maskunet::bn_act_fwd_stats", mutates_args=(), device_types="cuda")
def bn_act_fwd_stats(x: Tensor, residual: Tensor | None, gamma: Tensor, beta: Tensor, sums: Tensor, eps: float,
                     act: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """bn_act_fwd where ``sums`` f32 [2C] (per-channel sum, sum of squares of x over B*H*W) was already reduced
    by the convolution epilogue (mu_conv3x3_fwd): finalize + apply only, one HBM pass less."""
    M, C = _rows(x)
    if residual is not None:
        _rows(residual)
    _cuda(sums)
    y = torch.empty_like(x)
    f32 = dict(dtype=torch.float32, device=x.device)
    mean, rstd, a, b = (torch.empty((C,), **f32) for _ in range(4))
    with torch.cuda.device(x.device):
        _count(2)
        check(_L.mu_bn_act_fwd_stats(_p(x), _optp(residual), _p(gamma), _p(beta), eps, _p(y), _p(mean), _p(rstd),
                                     _p(a), _p(b), _p(sums), M, C, act, _code(x), _stream(x)), "mu_bn_act_fwd_stats")
    return y, mean, rstd, a, b


@bn_act_fwd_stats.register_fake
def _(x, residual, gamma, beta, sums, eps, act):
    C = x.shape[1]
    v = lambda: x.new_empty((C,), dtype=torch.float32)
    return torch.empty_like(x), v(), v(), v(), v()


def _bns_setup(ctx, inputs, output):
    x, residual, gamma, beta, sums, eps, act = inputs
    _bn_setup(ctx, (x, residual, gamma, beta, eps, act), output)


def _bns_backward(ctx, dy, *unused):
    dx, dr, dgamma, dbeta, _, _ = _bn_backward(ctx, dy)
    return dx, dr, dgamma, dbeta, None, None, None


bn_act_fwd_stats.register_autograd(_bns_backward, setup_context=_bns_setup)


# ------------------------------------------------------------------ K7: conv3x3 on tcgen05 (bf16 channels-last)
CONV_WIDTHS = (16, 32, 64, 128)


def conv3x3_shape_ok(B: int, Cin: int, Cout: int, H: int, W: int) -> bool:
    return (W in CONV_WIDTHS and H % (128 // W) == 0 and Cout % 64 == 0 and 64 <= Cout <= 512
            and ((Cin % 64 == 0 and 64 <= Cin <= 512) or (Cin % 8 == 0 and 8 <= Cin < 64)))


@torch.library.custom_op("maskunet::conv_prep_weights", mutates_args=(), device_types="cuda")
def conv_prep_weights(w: Tensor, with_wd: bool) -> Tuple[Tensor, Tensor]:
    """w f32 [Cout, Cin, kh, kw] -> wf bf16 [taps, Cout, Cin], wd bf16 [taps (reversed), Cin, Cout]."""
    _cuda(w)
    assert w.dtype == torch.float32 and w.dim() == 4
    Cout, Cin, kh, kw = w.shape
    taps = kh * kw
    wf = torch.empty((taps, Cout, (Cin + 63) // 64 * 64), dtype=torch.bfloat16, device=w.device)
    wd = torch.empty((taps, Cin, Cout) if with_wd else (0,), dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        _count(1)
        check(_L.mu_conv_prep_weights(_p(w), _p(wf), _p(wd) if with_wd else _optp(None), Cout, Cin, taps, _stream(w)),
              "mu_conv_prep_weights")
    return wf, wd


@conv_prep_weights.register_fake
def _(w, with_wd):
    Cout, Cin, kh, kw = w.shape
    return (w.new_empty((kh * kw, Cout, (Cin + 63) // 64 * 64), dtype=torch.bfloat16),
            w.new_empty((kh * kw, Cin, Cout) if with_wd else (0,), dtype=torch.bfloat16))


@torch.library.custom_op("maskunet::conv3x3_fwd", mutates_args=(), device_types="cuda")
def conv3x3_fwd(x: Tensor, wf: Tensor, want_stats: bool) -> Tuple[Tensor, Tensor]:
    """y = conv3x3(x) for bf16 channels-last x [B, Cin, H, W]; sums f32 [2 Cout] = per-channel (sum, sum of squares)
    of y from the epilogue (empty when not wanted)."""
    B, Cin, H, W = _nhwc(x)
    _cuda(wf)
    Cout = wf.shape[1]
    assert x.dtype == torch.bfloat16 and wf.shape == (9, Cout, (Cin + 63) // 64 * 64)
    y = _empty_cl(x, B, Cout, H, W)
    sums = torch.zeros((2 * Cout,), dtype=torch.float32, device=x.device) if want_stats else \
        torch.empty((0,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _timed("mu_conv3x3_fwd", (B, H, W, Cin, Cout)):
        _count(1)
        check(_L.mu_conv3x3_fwd(_p(x), _p(wf), _p(y), _p(sums) if want_stats else _optp(None), B, H, W, Cin, Cout,
                                MU_BF16, _stream(x)), "mu_conv3x3_fwd")
    return y, sums


@conv3x3_fwd.register_fake
def _(x, wf, want_stats):
    B, Cin, H, W = x.shape
    Cout = wf.shape[1]
    return _empty_cl(x, B, Cout, H, W), x.new_empty((2 * Cout if want_stats else 0,), dtype=torch.float32)


@torch.library.custom_op("maskunet::conv3x3_bwd_data", mutates_args=(), device_types="cuda")
def conv3x3_bwd_data(dy: Tensor, wd: Tensor) -> Tensor:
    B, Cout, H, W = _nhwc(dy)
    _cuda(wd)
    Cin = wd.shape[1]
    assert dy.dtype == torch.bfloat16 and wd.shape == (9, Cin, Cout)
    dx = _empty_cl(dy, B, Cin, H, W)
    with torch.cuda.device(dy.device), _timed("mu_conv3x3_bwd_data", (B, H, W, Cin, Cout)):
        _count(1)
        check(_L.mu_conv3x3_bwd_data(_p(dy), _p(wd), _p(dx), B, H, W, Cin, Cout, MU_BF16, _stream(dy)),
              "mu_conv3x3_bwd_data")
    return dx


@conv3x3_bwd_data.register_fake
def _(dy, wd):
    B, Cout, H, W = dy.shape
    return _empty_cl(dy, B, wd.shape[1], H, W)


@torch.library.custom_op("maskunet::conv3x3_bwd_weight", mutates_args=(), device_types="cuda")
def conv3x3_bwd_weight(x: Tensor, dy: Tensor) -> Tensor:
    """dw f32 [Cout, Cin, 3, 3]."""
    B, Cin, H, W = _nhwc(x)
    _, Cout, _, _ = _nhwc(dy)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16
    dw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=x.device)
    ws_bytes = int(_L.mu_conv3x3_workspace_bytes(Cin, Cout))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device), _timed("mu_conv3x3_bwd_weight", (B, H, W, Cin, Cout)):
        _count(2)
        check(_L.mu_conv3x3_bwd_weight(_p(x), _p(dy), _p(ws), ws_bytes, _p(dw), B, H, W, Cin, Cout, MU_BF16,
                                       _stream(x)), "mu_conv3x3_bwd_weight")
    return dw


@conv3x3_bwd_weight.register_fake
def _(x, dy):
    return x.new_empty((dy.shape[1], x.shape[1], 3, 3), dtype=torch.float32)


@torch.library.custom_op("maskunet::conv3x3", mutates_args=(), device_types="cuda")
def conv3x3(x: Tensor, weight: Tensor, want_stats: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """nn.Conv2d(Cin, Cout, 3, padding=1, bias=False) on our tcgen05 kernels.  x bf16 channels-last, weight the fp32
    parameter.  Returns (y, BatchNorm partial sums, wd) -- wd is the operand of the data gradient."""
    wf, wd = conv_prep_weights(weight.contiguous(), True)
    y, sums = conv3x3_fwd(x, wf, want_stats)
    return y, sums, wd


@conv3x3.register_fake
def _(x, weight, want_stats):
    B, Cin, H, W = x.shape
    Cout = weight.shape[0]
    return (_empty_cl(x, B, Cout, H, W), x.new_empty((2 * Cout if want_stats else 0,), dtype=torch.float32),
            x.new_empty((9, Cin, Cout), dtype=torch.bfloat16))


def _conv_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)   # unused outputs get None, not dense zero gradients
    x, weight, want_stats = inputs
    ctx.save_for_backward(x, output[2])
    ctx.x_needs_grad = x.requires_grad


def _conv_backward(ctx, dy, *unused):
    x, wd = ctx.saved_tensors
    dy = dy.contiguous(memory_format=torch.channels_last)
    dx = conv3x3_bwd_data(dy, wd) if ctx.needs_input_grad[0] else None
    dw = conv3x3_bwd_weight(x, dy) if ctx.needs_input_grad[1] else None
    return dx, dw, None


conv3x3.register_autograd(_conv_backward, setup_context=_conv_setup)


@torch.library.custom_op("maskunet::conv3x3_bwd_data_acc", mutates_args=("dx",), device_types="cuda")
def conv3x3_bwd_data_acc(dy: Tensor, wd: Tensor, dx: Tensor) -> None:
    """dx += conv(dy, wd): the TMA unit adds the output tiles into dx (bf16 reduce-add, every element exactly once)."""
    B, Cout, H, W = _nhwc(dy)
    _cuda(wd)
    Cin = wd.shape[1]
    assert dy.dtype == torch.bfloat16 and dx.dtype == torch.bfloat16 and wd.shape == (9, Cin, Cout)
    assert _nhwc(dx) == (B, Cin, H, W)
    with torch.cuda.device(dy.device), _timed("mu_conv3x3_bwd_data_acc", (B, H, W, Cin, Cout)):
        _count(1)
        check(_L.mu_conv3x3_bwd_data_acc(_p(dy), _p(wd), _p(dx), B, H, W, Cin, Cout, MU_BF16, _stream(dy)),
              "mu_conv3x3_bwd_data_acc")


@conv3x3_bwd_data_acc.register_fake
def _(dy, wd, dx):
    return None


def _addable(g: Optional[Tensor], like: Tensor) -> bool:
    """True when the gradient ``g`` of a second consumer can take our gradient of ``like`` in place."""
    return (g is not None and g.is_cuda and g.dtype == like.dtype and g.shape == like.shape and g.dim() == 4
            and g.is_contiguous(memory_format=torch.channels_last) and g._base is None and not g.requires_grad)


class _Conv3x3Skip(torch.autograd.Function):
    """conv3x3 of an activation that has a SECOND consumer (the residual branch of gelu(x + block(x)),
    ade_semantic.py:207).  Returns (y, sums, x_skip): ``x_skip`` is x itself, handed to the other consumer, so that
    both gradients of x arrive here and the data-gradient kernel adds its tiles straight into the other one --
    autograd's separate accumulation pass (one read of each gradient, one write) disappears."""

    @staticmethod
    def forward(ctx, x, weight, want_stats):
        wf, wd = conv_prep_weights(weight.contiguous(), True)
        y, sums = conv3x3_fwd(x, wf, want_stats)
        ctx.save_for_backward(x, wd)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(sums)
        return y, sums, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, _dsums, dskip):
        x, wd = ctx.saved_tensors
        dx = dw = None
        if dy is None:
            return (dskip if ctx.needs_input_grad[0] else None), None, None
        dy = dy.contiguous(memory_format=torch.channels_last)
        if ctx.needs_input_grad[0]:
            if _addable(dskip, x):
                conv3x3_bwd_data_acc(dy, wd, dskip)
                dx = dskip
            else:
                dx = conv3x3_bwd_data(dy, wd)
                if dskip is not None:
                    dx = dx + dskip
        if ctx.needs_input_grad[1]:
            dw = conv3x3_bwd_weight(x, dy)
        return dx, dw, None


def conv3x3_skip(x: Tensor, weight: Tensor, want_stats: bool) -> Tuple[Tensor, Tensor, Tensor]:
    return _Conv3x3Skip.apply(x, weight, want_stats)


# ------------------------------------------------------------------ K12: 1x1 convolution heads (class-padded outputs)
HEAD_PADS = (32, 64, 128, 160, 256)


def pad_channels(c: int) -> int:
    """Padded channel count of a 1x1 head output (TMA needs a 16-byte row pitch; 150 classes -> 160)."""
    for n in HEAD_PADS:
        if c <= n:
            return n
    raise ValueError(f"1x1 head with {c} output channels is not supported (max {HEAD_PADS[-1]})")


def conv1x1_shape_ok(Cin: int, Cout: int, H: int, W: int) -> bool:
    return (W in CONV_WIDTHS and H % (128 // W) == 0 and Cin in (64, 128, 256) and 1 <= Cout <= HEAD_PADS[-1])


@torch.library.custom_op("maskunet::column_sums", mutates_args=(), device_types="cuda")
def column_sums(x: Tensor) -> Tensor:
    """x channels-last [B, C, H, W] -> f32 [2C]: per-channel sum and sum of squares."""
    M, C = _rows(x)
    sums = torch.empty((2 * C,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _count(1)
        check(_L.mu_column_sums(_p(x), _p(sums), M, C, _code(x), _stream(x)), "mu_column_sums")
    return sums


@column_sums.register_fake
def _(x):
    return x.new_empty((2 * x.shape[1],), dtype=torch.float32)


@torch.library.custom_op("maskunet::conv1x1", mutates_args=(), device_types="cuda")
def conv1x1(x: Tensor, weight: Tensor, bias: Tensor | None, n_pad: int, want_stats: bool = False
            ) -> Tuple[Tensor, Tensor, Tensor]:
    """nn.Conv2d(Cin, Cout, 1) on tcgen05: x bf16 channels-last [B, Cin, H, W], weight f32 [Cout, Cin, 1, 1],
    bias f32 [Cout].  Returns (y [B, n_pad, H, W] with channels >= Cout zero, wd = data-gradient operand, sums f32
    [2 n_pad] = per-channel (sum, sum of squares) of y from the epilogue -- the statistics pass of the BatchNorm2d that
    follows the head -- or an empty tensor when not wanted)."""
    B, Cin, H, W = _nhwc(x)
    _cuda(weight)
    Cout = weight.shape[0]
    assert x.dtype == torch.bfloat16 and weight.dtype == torch.float32 and weight.shape[1] == Cin and n_pad >= Cout
    dev = x.device
    wf = torch.empty((n_pad, (Cin + 63) // 64 * 64), dtype=torch.bfloat16, device=dev)
    wd = torch.empty((Cin, (n_pad + 63) // 64 * 64), dtype=torch.bfloat16, device=dev)
    bias_p = torch.empty((n_pad,), dtype=torch.float32, device=dev)
    y = _empty_cl(x, B, n_pad, H, W)
    sums = torch.zeros((2 * n_pad,), dtype=torch.float32, device=dev) if want_stats else \
        torch.empty((0,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _timed("mu_conv1x1_fwd", (B, H, W, Cin, n_pad)):
        _count(2)
        check(_L.mu_conv1x1_prep(_p(weight.contiguous()), _optp(bias), _p(wf), _p(wd), _p(bias_p), Cout, Cin, n_pad,
                                 _stream(x)), "mu_conv1x1_prep")
        if want_stats:
            check(_L.mu_conv1x1_fwd_stats(_p(x), _p(wf), _p(bias_p), _p(y), _p(sums), B, H, W, Cin, n_pad, MU_BF16,
                                          _stream(x)), "mu_conv1x1_fwd_stats")
        else:
            check(_L.mu_conv1x1_fwd(_p(x), _p(wf), _p(bias_p), _p(y), B, H, W, Cin, n_pad, MU_BF16, _stream(x)),
                  "mu_conv1x1_fwd")
    return y, wd, sums


@conv1x1.register_fake
def _(x, weight, bias, n_pad, want_stats=False):
    B, Cin, H, W = x.shape
    return (_empty_cl(x, B, n_pad, H, W), x.new_empty((Cin, (n_pad + 63) // 64 * 64), dtype=torch.bfloat16),
            x.new_empty((2 * n_pad if want_stats else 0,), dtype=torch.float32))


@torch.library.custom_op("maskunet::conv1x1_bwd", mutates_args=(), device_types="cuda")
def conv1x1_bwd(x: Tensor, dy: Tensor, wd: Tensor, want_dx: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (dx [B, Cin, H, W] (empty when not wanted), dw f32 [n_pad, Cin], db f32 [n_pad])."""
    B, Cin, H, W = _nhwc(x)
    _, n_pad, _, _ = _nhwc(dy)
    dev = x.device
    dx = _empty_cl(x, B, Cin, H, W) if want_dx else x.new_empty((0,))
    dw = torch.empty((n_pad, Cin), dtype=torch.float32, device=dev)
    ws_bytes = max(int(_L.mu_conv1x1_workspace_bytes(Cin, n_pad)), 2 * n_pad * 4)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    db = torch.empty((n_pad,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _timed("mu_conv1x1_bwd", (B, H, W, Cin, n_pad)):
        # data gradient, weight gradient + its finish kernel; the bias gradient rides in the weight-gradient GEMM (64 input
        # channels, free-running mode) or costs one more kernel
        folded = Cin == 64 and not _L.mu_get_deterministic()
        _count((3 if want_dx else 2) + (0 if folded else 1))
        if want_dx:
            check(_L.mu_conv1x1_bwd_data(_p(dy), _p(wd), _p(dx), B, H, W, Cin, n_pad, MU_BF16, _stream(x)),
                  "mu_conv1x1_bwd_data")
        check(_L.mu_conv1x1_bwd_weight_bias(_p(x), _p(dy), _p(ws), ws_bytes, _p(dw), _p(db), B, H, W, Cin, n_pad, MU_BF16,
                                            _stream(x)), "mu_conv1x1_bwd_weight_bias")
    return dx, dw, db


@conv1x1_bwd.register_fake
def _(x, dy, wd, want_dx):
    n_pad = dy.shape[1]
    return (torch.empty_like(x) if want_dx else x.new_empty((0,)), x.new_empty((n_pad, x.shape[1]), dtype=torch.float32),
            x.new_empty((n_pad,), dtype=torch.float32))


def _c1_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    x, weight, bias, n_pad = inputs[:4]
    ctx.save_for_backward(x, output[1])
    ctx.cout = weight.shape[0]
    ctx.has_bias = bias is not None


def _c1_backward(ctx, dy, *unused):
    x, wd = ctx.saved_tensors
    dy = dy.contiguous(memory_format=torch.channels_last)
    dx, dw, db = conv1x1_bwd(x, dy, wd, bool(ctx.needs_input_grad[0]))
    cout = ctx.cout
    return (dx if ctx.needs_input_grad[0] else None, dw[:cout].reshape(cout, x.shape[1], 1, 1).contiguous(),
            db[:cout].contiguous() if ctx.has_bias else None, None, None)


conv1x1.register_autograd(_c1_backward, setup_context=_c1_setup)


# ------------------------------------------------------------------ K9: MaxPool2d(2), channels-last
def _nhwc(x: Tensor):
    if not (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        raise RuntimeError("this maskunet op takes 4-D channels-last tensors")
    return x.shape[0], x.shape[1], x.shape[2], x.shape[3]          # B, C, H, W


def _empty_cl(x: Tensor, B, C, H, W):
    return torch.empty((B, C, H, W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)


@torch.library.custom_op("maskunet::maxpool2", mutates_args=(), device_types="cuda")
def maxpool2(x: Tensor) -> Tensor:
    B, C, H, W = _nhwc(x)
    y = _empty_cl(x, B, C, H // 2, W // 2)
    with torch.cuda.device(x.device):
        _count(1)
        check(_L.mu_maxpool2(_p(x), _optp(None), _p(y), B, H, W, C, 0, _code(x), _stream(x)), "mu_maxpool2")
    return y


@maxpool2.register_fake
def _(x):
    B, C, H, W = x.shape
    return _empty_cl(x, B, C, H // 2, W // 2)


@torch.library.custom_op("maskunet::maxpool2_bwd", mutates_args=(), device_types="cuda")
def maxpool2_bwd(x: Tensor, dy: Tensor) -> Tensor:
    B, C, H, W = _nhwc(x)
    _nhwc(dy)
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _count(1)
        check(_L.mu_maxpool2(_p(x), _p(dy), _p(dx), B, H, W, C, 1, _code(x), _stream(x)), "mu_maxpool2")
    return dx


@maxpool2_bwd.register_fake
def _(x, dy):
    return torch.empty_like(x)


maxpool2.register_autograd(
    lambda ctx, g: maxpool2_bwd(ctx.saved_tensors[0], g.contiguous(memory_format=torch.channels_last)),
    setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


@torch.library.custom_op("maskunet::maxpool2_bwd_acc", mutates_args=("dx",), device_types="cuda")
def maxpool2_bwd_acc(x: Tensor, dy: Tensor, dx: Tensor) -> None:
    """dx += maxpool2 gradient (dx holds the skip connection's gradient of x)."""
    B, C, H, W = _nhwc(x)
    _nhwc(dy)
    assert _nhwc(dx) == (B, C, H, W) and dx.dtype == x.dtype
    with torch.cuda.device(x.device):
        _count(1)
        check(_L.mu_maxpool2(_p(x), _p(dy), _p(dx), B, H, W, C, 2, _code(x), _stream(x)), "mu_maxpool2")


@maxpool2_bwd_acc.register_fake
def _(x, dy, dx):
    return None


class _MaxPool2Skip(torch.autograd.Function):
    """MaxPool2d(2) of an activation that is also a U-Net skip connection (ade_semantic.py:301-312: x1, x2, x3 feed a
    DownSample and, later, an UpSample's concat).  Returns (pooled, x_skip); the pooling gradient is added into the
    skip gradient in place (see _Conv3x3Skip)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        ctx.set_materialize_grads(False)
        return maxpool2(x), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        (x,) = ctx.saved_tensors
        if dy is None:
            return dskip
        dy = dy.contiguous(memory_format=torch.channels_last)
        if _addable(dskip, x):
            maxpool2_bwd_acc(x, dy, dskip)
            return dskip
        dx = maxpool2_bwd(x, dy)
        return dx if dskip is None else dx + dskip


def maxpool2_skip(x: Tensor) -> Tuple[Tensor, Tensor]:
    return _MaxPool2Skip.apply(x)


# ------------------------------------------------------------------ K10: bilinear x2 (align_corners) + concat
@torch.library.custom_op("maskunet::upsample_concat", mutates_args=(), device_types="cuda")
def upsample_concat(skip: Tensor, x: Tensor) -> Tensor:
    """cat([skip, Upsample(2, bilinear, align_corners=True)(x)], dim=1) on channels-last tensors."""
    B, Cx, H, W = _nhwc(x)
    Bs, Cs, H2, W2 = _nhwc(skip)
    assert (Bs, H2, W2) == (B, 2 * H, 2 * W) and skip.dtype == x.dtype
    out = _empty_cl(x, B, Cs + Cx, H2, W2)
    with torch.cuda.device(x.device):
        _count(1)
        check(_L.mu_upsample_concat_fwd(_p(skip), _p(x), _p(out), B, H, W, Cs, Cx, _code(x), _stream(x)),
              "mu_upsample_concat_fwd")
    return out


@upsample_concat.register_fake
def _(skip, x):
    B, Cx, H, W = x.shape
    return _empty_cl(x, B, skip.shape[1] + Cx, 2 * H, 2 * W)


@torch.library.custom_op("maskunet::upsample_concat_bwd", mutates_args=(), device_types="cuda")
def upsample_concat_bwd(dout: Tensor, cs: int) -> Tuple[Tensor, Tensor]:
    B, Ct, H2, W2 = _nhwc(dout)
    cx, H, W = Ct - cs, H2 // 2, W2 // 2
    dskip = _empty_cl(dout, B, cs, H2, W2)
    dx = _empty_cl(dout, B, cx, H, W)
    with torch.cuda.device(dout.device):
        _count(1)
        check(_L.mu_upsample_concat_bwd(_p(dout), _p(dskip), _p(dx), B, H, W, cs, cx, _code(dout), _stream(dout)),
              "mu_upsample_concat_bwd")
    return dskip, dx


@upsample_concat_bwd.register_fake
def _(dout, cs):
    B, Ct, H2, W2 = dout.shape
    return _empty_cl(dout, B, cs, H2, W2), _empty_cl(dout, B, Ct - cs, H2 // 2, W2 // 2)


def _upcat_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)   # unused outputs get None, not dense zero gradients
    ctx.cs = inputs[0].shape[1]


upsample_concat.register_autograd(
    lambda ctx, g: upsample_concat_bwd(g.contiguous(memory_format=torch.channels_last), ctx.cs),
    setup_context=_upcat_setup)


# ------------------------------------------------------------------ K11: LayerNorm over (C, H, W) per sample
@torch.library.custom_op("maskunet::sample_layernorm", mutates_args=(), device_types="cuda")
def sample_layernorm(x: Tensor, gamma_nhwc: Tensor, beta_nhwc: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """x channels-last [B, C, H, W]; gamma / beta f32 [H, W, C] (the [C, H, W] parameters permuted)."""
    B, C, H, W = _nhwc(x)
    _cuda(gamma_nhwc, beta_nhwc)
    L = C * H * W
    y = torch.empty_like(x)
    f32 = dict(dtype=torch.float32, device=x.device)
    mean, rstd, sums = torch.empty((B,), **f32), torch.empty((B,), **f32), torch.empty((2 * B,), **f32)
    with torch.cuda.device(x.device):
        _count(3)
        check(_L.mu_sample_layernorm_fwd(_p(x), _p(gamma_nhwc), _p(beta_nhwc), eps, _p(y), _p(mean), _p(rstd),
                                         _p(sums), B, L, _code(x), _stream(x)), "mu_sample_layernorm_fwd")
    return y, mean, rstd


@sample_layernorm.register_fake
def _(x, gamma_nhwc, beta_nhwc, eps):
    v = x.new_empty((x.shape[0],), dtype=torch.float32)
    return torch.empty_like(x), v, torch.empty_like(v)


@torch.library.custom_op("maskunet::sample_layernorm_bwd", mutates_args=(), device_types="cuda")
def sample_layernorm_bwd(dy: Tensor, x: Tensor, gamma_nhwc: Tensor, mean: Tensor, rstd: Tensor
                         ) -> Tuple[Tensor, Tensor, Tensor]:
    B, C, H, W = _nhwc(x)
    _nhwc(dy)
    L = C * H * W
    dx = torch.empty_like(x)
    dgamma, dbeta = torch.empty_like(gamma_nhwc), torch.empty_like(gamma_nhwc)
    sums = torch.empty((2 * B,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _count(2)
        check(_L.mu_sample_layernorm_bwd(_p(dy), _p(x), _p(gamma_nhwc), _p(mean), _p(rstd), _p(sums), _p(dx),
                                         _p(dgamma), _p(dbeta), B, L, _code(x), _stream(x)),
              "mu_sample_layernorm_bwd")
    return dx, dgamma, dbeta


@sample_layernorm_bwd.register_fake
def _(dy, x, gamma_nhwc, mean, rstd):
    return torch.empty_like(x), torch.empty_like(gamma_nhwc), torch.empty_like(gamma_nhwc)


def _sln_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)   # unused outputs get None, not dense zero gradients
    x, gamma, beta, eps = inputs
    ctx.save_for_backward(x, gamma, output[1], output[2])


def _sln_backward(ctx, dy, *unused):
    x, gamma, mean, rstd = ctx.saved_tensors
    dx, dg, db = sample_layernorm_bwd(dy.contiguous(memory_format=torch.channels_last), x, gamma, mean, rstd)
    return dx, dg, db, None


sample_layernorm.register_autograd(_sln_backward, setup_context=_sln_setup)


# ------------------------------------------------------------------ A14: fused cross-entropy (mean) + gradient
@torch.library.custom_op("maskunet::cross_entropy_fused", mutates_args=(), device_types="cuda")
def cross_entropy_fused(logits: Tensor, labels: Tensor, ignore_index: int, n_classes: int = -1,
                        valid_count: Tensor | None = None) -> Tuple[Tensor, Tensor]:
    """logits channels-last [B, P, H, W], labels int64 [B, H, W] -> (mean loss [1] f32, dlogits like logits).
    n_classes > 0: only the first n_classes of the P channels are classes (the class-padded output of the 1x1
    head); the pad channels of dlogits are zero.
    valid_count f32 [1]: the normaliser of the loss sum when it is not this call's own valid-pixel count (micro-batches
    and data-parallel ranks normalise by the count of the whole global batch, see train.py).
    A label outside [0, C) that is not ``ignore_index`` (nn.CrossEntropyLoss: device assert) makes the loss and that
    row's gradient NaN; with every row ignored the loss is NaN and the gradient zero, as in torch."""
    B, P, H, W = _nhwc(logits)
    C = n_classes if n_classes > 0 else P
    _cuda(labels)
    assert labels.dtype == torch.int64 and labels.shape == (B, H, W) and C <= P
    dlogits = torch.empty_like(logits)
    loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
    if valid_count is None:
        count = (labels != ignore_index).sum().to(torch.float32).reshape(1)
    else:
        _cuda(valid_count)
        assert valid_count.dtype == torch.float32 and valid_count.numel() == 1
        count = valid_count
    with torch.cuda.device(logits.device):
        _count(1)
        check(_L.mu_cross_entropy_fused(_p(logits), _p(labels), _p(count), ignore_index, _p(dlogits), _p(loss),
                                        B * H * W, C, P, _code(logits), _stream(logits)), "mu_cross_entropy_fused")
    return loss, dlogits


@cross_entropy_fused.register_fake
def _(logits, labels, ignore_index, n_classes=-1, valid_count=None):
    return logits.new_empty((1,), dtype=torch.float32), torch.empty_like(logits)


def _ce_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)   # unused outputs get None, not dense zero gradients
    ctx.save_for_backward(output[1])


def _ce_backward(ctx, dloss, *unused):
    (dlogits,) = ctx.saved_tensors
    return dlogits * dloss.to(dlogits.dtype), None, None, None, None


cross_entropy_fused.register_autograd(_ce_backward, setup_context=_ce_setup)


# ------------------------------------------------------------------ logit post-processing
def _class_rows(logits: Tensor):
    """(tensor to pass, pitch): channels-last logits, possibly the [:, :C] view of a class-padded buffer."""
    B, C, H, W = logits.shape
    st = logits.stride()
    if logits.is_contiguous(memory_format=torch.channels_last):
        return logits, C
    if st[1] == 1 and st[3] >= C and st[2] == W * st[3] and st[0] == H * W * st[3]:
        return logits, st[3]                                  # rows of pitch st[3], first C entries are the classes
    return logits.contiguous(memory_format=torch.channels_last), C


@torch.library.custom_op("maskunet::argmax_iou", mutates_args=(), device_types="cuda")
def argmax_iou(logits: Tensor, labels: Tensor | None, pitch: int, smooth: float) -> Tuple[Tensor, Tensor, Tensor]:
    """logits [B, C, H, W] whose memory is rows of `pitch` >= C channels -> (pred int64 [B, H, W], mean IoU f32 [1]
    (NaN without labels), hist int32 [3, C]: predicted / labelled / matching pixels per class)."""
    B, C, H, W = logits.shape
    dev = logits.device
    pred = torch.empty((B, H, W), dtype=torch.int64, device=dev)
    hist = torch.zeros((3, C), dtype=torch.int32, device=dev)
    miou = torch.full((1,), float("nan"), dtype=torch.float32, device=dev)
    if labels is not None:
        _cuda(labels)
        assert labels.dtype == torch.int64 and labels.shape == (B, H, W)
    with torch.cuda.device(dev):
        _count(2 if labels is not None else 1)
        check(_L.mu_argmax_iou(_p(logits), _optp(labels), _p(pred), _p(hist), _p(miou) if labels is not None else _optp(None),
                               B * H * W, C, pitch, smooth, _code(logits), _stream(logits)), "mu_argmax_iou")
    return pred, miou, hist


@argmax_iou.register_fake
def _(logits, labels, pitch, smooth):
    B, C, H, W = logits.shape
    return (logits.new_empty((B, H, W), dtype=torch.int64), logits.new_empty((1,), dtype=torch.float32),
            logits.new_empty((3, C), dtype=torch.int32))


def segmentation_argmax(y_pred: Tensor) -> Tensor:
    """argmax(softmax(y_pred / 0.5, dim=1), dim=1) (ade_semantic.py:130-131) on the device, one pass."""
    t, pitch = _class_rows(y_pred.detach())
    return argmax_iou(t, None, pitch, 0.0)[0]


def mean_iou(y_pred: Tensor, y_true: Tensor, num_classes: int, smooth: float = 1e-6) -> Tensor:
    """Drop-in for the reference's mean_iou (ade_semantic.py:128-146): same signature, same value, no host syncs."""
    if y_pred.shape[1] != num_classes:
        raise ValueError("mean_iou: y_pred must have num_classes channels")
    t, pitch = _class_rows(y_pred.detach())
    return argmax_iou(t, y_true.contiguous(), pitch, smooth)[1][0]


# ------------------------------------------------------------------ the module-level op (A3-A9 of SURVEY.md 8(a))
@torch.library.custom_op("maskunet::mask_attention", mutates_args=(), device_types="cuda")
def mask_attention(x: Tensor, w_qkv: Tensor, b_qkv: Tensor, gamma: Tensor, beta: Tensor, keep_rank: Tensor,
                   keep_idx: Tensor, n_keep: Tensor, eps: float, token_major: bool, view_out: bool = False
                   ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """y [B, N, C] = LN_C(softmax(QK^T/sqrt(C) + mask) V + tokens); also returns what backward needs.
    ``view_out`` (bf16 token-major, N % C == 0): y comes back as the channels-last memory of the module's re-viewed
    result, y[b, p, c] = result[b].flat[c N + p] (ade_semantic.py:190), and the backward takes its gradient likewise --
    the transpose passes between the module and the convolutions around it happen inside the LayerNorm kernels."""
    q, kc, vc = qkv_project(x, w_qkv, b_qkv, keep_rank, n_keep, token_major)
    o, lse = attn_fwd(q, kc, vc, n_keep)
    y, mean, rstd = residual_ln_fwd(o, x, gamma, beta, eps, token_major, view_out)
    return y, q, kc, vc, o, lse, mean, rstd


@mask_attention.register_fake
def _(x, w_qkv, b_qkv, gamma, beta, keep_rank, keep_idx, n_keep, eps, token_major, view_out=False):
    B, C, N = _bnc(x, token_major)
    f32 = dict(dtype=torch.float32)
    return (x.new_empty((B, N, C)), x.new_empty((B, N, C)), x.new_empty((B, nkp_of(N), C)),
            x.new_empty((B, nkp_of(N), C)), x.new_empty((B, N, C)), x.new_empty((B, N), **f32),
            x.new_empty((B, N), **f32), x.new_empty((B, N), **f32))


@torch.library.custom_op("maskunet::mask_attention_bwd", mutates_args=(), device_types="cuda")
def mask_attention_bwd(dy: Tensor, x: Tensor, w_qkv: Tensor, gamma: Tensor, keep_idx: Tensor, n_keep: Tensor,
                       q: Tensor, kc: Tensor, vc: Tensor, o: Tensor, lse: Tensor, mean: Tensor, rstd: Tensor,
                       token_major: bool, view_out: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    dz, delta, dgamma, dbeta = residual_ln_bwd(dy, o, x, mean, rstd, gamma, token_major, view_out)
    dq, dk, dv = attn_bwd(q, kc, vc, n_keep, keep_idx, dz, lse, delta)
    dx, dw, db = qkv_project_bwd(x, dz, dq, dk, dv, w_qkv, token_major)
    return dx, dw, db, dgamma, dbeta


@mask_attention_bwd.register_fake
def _(dy, x, w_qkv, gamma, keep_idx, n_keep, q, kc, vc, o, lse, mean, rstd, token_major, view_out=False):
    C = w_qkv.shape[1]
    f32 = dict(dtype=torch.float32)
    return (torch.empty_like(x), x.new_empty((3 * C, C), **f32), x.new_empty((3 * C,), **f32),
            x.new_empty((C,), **f32), x.new_empty((C,), **f32))


def _ma_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)   # unused outputs get None, not dense zero gradients
    x, w_qkv, b_qkv, gamma, beta, keep_rank, keep_idx, n_keep, eps, token_major = inputs[:10]
    y, q, kc, vc, o, lse, mean, rstd = output
    ctx.token_major = token_major
    ctx.view_out = bool(inputs[10]) if len(inputs) > 10 else False
    ctx.save_for_backward(x, w_qkv, gamma, keep_idx, n_keep, q, kc, vc, o, lse, mean, rstd)


def _ma_backward(ctx, dy, *unused):
    x, w_qkv, gamma, keep_idx, n_keep, q, kc, vc, o, lse, mean, rstd = ctx.saved_tensors
    dx, dw, db, dgamma, dbeta = mask_attention_bwd(dy.contiguous(), x, w_qkv, gamma, keep_idx, n_keep,
                                                   q, kc, vc, o, lse, mean, rstd, ctx.token_major, ctx.view_out)
    return dx, dw, db, dgamma, dbeta, None, None, None, None, None, None


mask_attention.register_autograd(_ma_backward, setup_context=_ma_setup)
