"""Drop-in nn.Modules for the reference's model block.

Same class names, constructor and forward signatures, attribute names and
state_dict keys as /root/reference/code/ade20k/ade_semantic.py:152-314 (and the
three-output variant /root/reference/code/cityscapes/city_instance.py:216-276),
so a script can replace its inline classes with

    from maskunet_b200 import Mask2FormerAttention, ConvBlock, DownSample, UpSample, UNet

and a reference checkpoint loads unchanged.  The Mask Attention Module runs on
the hand-written sm_100a kernels behind ``maskunet::mask_attention``; it needs
CUDA tensors and has no CPU path.  Extra switches are keyword-only and default
to the reference's behaviour.
"""
from __future__ import annotations

import os
import threading
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class Mask2FormerAttention(nn.Module):
    """Single-head masked self-attention over H*W tokens + residual + LayerNorm.

    Reference: ade_semantic.py:152-190.  Semantics kept exactly:
      * tokens are x[b, :, n]; Q/K/V = Linear with bias; scores / sqrt(C)
      * mask: per sample, per KEY, shared by all queries, drawn once by the same
        ``torch.randint(0, 2, (B, H, W), device=x.device)`` call and cached in ``self.mask`` as the
        expanded [B, N, N] 0/-inf view; regenerated only when None or N changes; not in state_dict
      * output is the [B, N, C] result re-viewed (not permuted back) as [B, C, H, W]

    Keyword-only extras: ``mask_mode='cached'|'resample'`` ('resample' draws a new mask every forward,
    which is what the reference does under multi-GPU nn.DataParallel);
    ``compute_dtype`` (None = follow the input: float32 -> fp32 validation kernels, bfloat16 -> tcgen05).
    """

    def __init__(self, channels: int, size: int, *, mask_mode: str = "cached",
                 compute_dtype: Optional[torch.dtype] = None):
        super().__init__()
        if mask_mode not in ("cached", "resample"):
            raise ValueError("mask_mode must be 'cached' or 'resample'")
        self.channels = channels
        self.size = size
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.mask = None
        self.norm = nn.LayerNorm([channels])
        self.mask_mode = mask_mode
        self.compute_dtype = compute_dtype
        self._compaction = None  # (mask object it was derived from, keep_bits, n_keep, keep_idx, keep_rank)

    # -- mask handling ---------------------------------------------------------------------------
    def _draw_mask(self, batch: int, height: int, width: int, device) -> None:
        bits = torch.randint(0, 2, (batch, height, width), device=device)          # :178, same RNG use
        flat = bits.view(batch, -1)
        comp = ops.mask_binarize(flat)                                               # :179-180 on device
        zero = torch.tensor(0.0, device=device)
        ninf = torch.tensor(-float("inf"), device=device)
        bias = torch.where(comp[3] >= 0, zero, ninf)                                 # keep_rank >= 0 <=> kept
        self.mask = bias.unsqueeze(1).expand(-1, height * width, -1)                 # :181, stride (N, 0, 1)
        self._compaction = (self.mask,) + tuple(comp)

    def _compaction_for(self, mask: torch.Tensor):
        """keep_bits / n_keep / keep_idx / keep_rank for the current ``self.mask`` (which tests may inject)."""
        if self._compaction is None or self._compaction[0] is not mask:
            keep = (mask[:, 0, :] == 0).to(torch.int64).contiguous()
            self._compaction = (mask,) + tuple(ops.mask_binarize(keep))
        return self._compaction[1:]

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        batch, channels, height, width = x.size()
        if channels != self.channels:
            raise ValueError("Input channel size does not match initialized channel size.")
        if not x.is_cuda:
            raise RuntimeError("maskunet_b200.Mask2FormerAttention runs on CUDA (sm_100a) only; "
                               "there is no CPU fallback")
        n_tok = height * width
        if self.mask_mode == "resample" or self.mask is None or self.mask.size(-1) != n_tok:
            self._draw_mask(batch, height, width, x.device)
        if self.mask.size(0) != batch:
            raise RuntimeError(f"The size of tensor a ({batch}) must match the size of tensor b "
                               f"({self.mask.size(0)}) at non-singleton dimension 0 (cached mask batch)")
        _, n_keep, keep_idx, keep_rank = self._compaction_for(self.mask)

        dtype = self.compute_dtype or x.dtype
        if dtype not in (torch.float32, torch.bfloat16):
            dtype = torch.float32
        channels_last_in = x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)
        if dtype == torch.bfloat16:
            # tensor-core path: token-major [B, N, C] activations.  Channels-last memory IS that layout (the
            # reference's permuted view of :168); NCHW input goes through the coalesced transpose kernel.
            token_major = True
            if channels_last_in:
                tokens = x.permute(0, 2, 3, 1).reshape(batch, n_tok, channels).to(dtype)
            else:
                tokens = ops.transpose(x.contiguous().view(batch, channels, n_tok).to(dtype))
        else:
            token_major = False
            tokens = x.contiguous().view(batch, channels, n_tok).to(dtype)
        def pack():
            return (torch.cat([self.query.weight, self.key.weight, self.value.weight], dim=0).float(),
                    torch.cat([self.query.bias, self.key.bias, self.value.bias], dim=0).float())
        if torch.is_grad_enabled():
            w_qkv, b_qkv = pack()
        else:       # inference: packed once per parameter version (see _inference_cache)
            w_qkv, b_qkv = _inference_cache(self, "_mu_qkv_operand", (self.query.weight, self.key.weight,
                                                                    self.value.weight, self.query.bias,
                                                                    self.key.bias, self.value.bias), pack)
        # channels-last in, channels-last out: the re-view of :190 (the [B, N, C] buffer read as [B, C, H, W]) stored
        # channels-last is a transpose of the [C, N] view of each sample; the LayerNorm kernels do it on the way out
        # (and on the way in, for the gradient) when the geometry allows, a separate transpose pass otherwise
        view_out = (channels_last_in and token_major and channels in (64, 128, 256) and n_tok % channels == 0
                    and os.environ.get("MASKUNET_LN_VIEW", "1") != "0")
        y = ops.mask_attention(tokens, w_qkv, b_qkv, self.norm.weight.float(), self.norm.bias.float(),
                               keep_rank, keep_idx, n_keep, self.norm.eps, token_major, view_out)[0]
        if channels_last_in:
            if not view_out:
                y = ops.transpose(y.view(batch, channels, n_tok))
            return y.view(batch, height, width, channels).permute(0, 3, 1, 2)
        return y.view(batch, channels, height, width)                               # :190


def _fast_layout(x: torch.Tensor, channel_multiple: int = 8) -> bool:
    """True when x is a CUDA channels-last activation our NHWC kernels take."""
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16)
            and x.shape[1] % channel_multiple == 0 and x.is_contiguous(memory_format=torch.channels_last))


# BatchNorm `num_batches_tracked += 1` is one tiny kernel per layer; UNet.forward collects the counters of its own
# forward in a per-thread list and bumps them with a single multi-tensor add.  Thread-local and saved / restored around
# the forward, so nn.DataParallel replicas (threads) and nested or re-entrant UNets never see each other's list; outside
# a UNet forward there is no list and each layer updates its own counter.
_TLS = threading.local()


def _nbt_list() -> Optional[list]:
    return getattr(_TLS, "nbt", None)


def _bn_eval_affine(x: torch.Tensor, bn: nn.BatchNorm2d, n_pad: int,
                    pre_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Eval-mode BatchNorm2d as y = a x + b on a tensor whose trailing ``n_pad`` channels are zero padding
    (differentiable w.r.t. x, weight, bias and ``pre_bias``; the pad channels get a = b = 0).  The affine is
    evaluated in fp32 and rounded once, as torch's own eval BatchNorm does.  ``pre_bias``: a per-channel bias the
    producing convolution left out (BN(x + pre_bias))."""
    a = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
    shift = bn.running_mean.float() if pre_bias is None else bn.running_mean.float() - pre_bias.float()
    b = bn.bias.float() - shift * a
    a, b = F.pad(a, (0, n_pad)), F.pad(b, (0, n_pad))
    return (x.float() * a.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)).to(x.dtype)


class _AttachZeroGrad(torch.autograd.Function):
    """y unchanged; ``param`` receives an all-zero gradient.  A convolution bias in front of a training-mode BatchNorm
    cancels in the forward and has an exactly-zero gradient: the kernels skip it, and this node still hands the
    optimiser the zero gradient the reference's autograd produces (AdamW decays a parameter whose gradient is zero,
    and skips one whose gradient is None)."""

    @staticmethod
    def forward(ctx, y, param):
        ctx.shape, ctx.dtype, ctx.device = param.shape, param.dtype, param.device
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        return g, torch.zeros(ctx.shape, dtype=ctx.dtype, device=ctx.device)


def fused_bn_act(x: torch.Tensor, bn: nn.BatchNorm2d, act: int, residual: Optional[torch.Tensor] = None,
                 sums: Optional[torch.Tensor] = None, pre_bias: Optional[torch.Tensor] = None):
    """act(BatchNorm2d(x) [+ residual]) -- the BN / GELU / ReLU / residual chains of ade_semantic.py:198-210,
    :219, :240 and :283-287.  ``sums`` = per-channel (sum, sum of squares) of x already reduced by the producer
    (the epilogue of our convolution kernel), which removes the statistics pass.  ``pre_bias``: the bias of the
    producing convolution when the kernel left it out -- the result is act(BN(x + pre_bias) [+ residual]): under batch
    statistics the bias cancels (only the running mean sees it), in eval mode it folds into the affine.

    Channels-last CUDA activations (the bf16 production layout) run on our fused sm_100a kernels
    (csrc/bn_act.cu: one statistics pass, one apply pass; backward one reduce + one apply pass).  Any other
    layout (the NCHW fp32 validation configuration) is evaluated with the stock torch CUDA ops in the
    reference's own order.
    """
    batch_stats = bn.training or not bn.track_running_stats      # nn.BatchNorm2d: no running statistics -> batch ones
    fused = (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16)
             and x.is_contiguous(memory_format=torch.channels_last) and bn.affine
             and (batch_stats or not torch.is_grad_enabled()))
    if not fused:
        n_pad = x.shape[1] - bn.num_features
        if (n_pad > 0 or pre_bias is not None) and not batch_stats and bn.affine:
            # class-padded head output in eval() with autograd on (frozen-BN fine-tuning, gradient checks): the
            # inference affine form in plain torch ops, pad channels stay zero (a = b = 0, act(0) = 0)
            y = _bn_eval_affine(x, bn, n_pad, pre_bias)
        else:
            if n_pad > 0:
                raise RuntimeError("fused_bn_act: a class-padded activation needs an affine BatchNorm2d "
                                   f"({x.shape[1]} channels against {bn.num_features} features)")
            if pre_bias is not None:
                x = x + pre_bias.to(x.dtype).view(1, -1, 1, 1)
            y = bn(x)
        if residual is not None:
            y = residual + y
        if act == ops.ACT_GELU:
            return F.gelu(y)
        if act == ops.ACT_RELU:
            return F.relu(y)
        return y
    if residual is not None:
        residual = residual.to(x.dtype).contiguous(memory_format=torch.channels_last)
    gamma, beta = bn.weight.float(), bn.bias.float()
    n_feat, n_pad = bn.num_features, x.shape[1] - bn.num_features
    if n_pad > 0:       # class-padded output of our 1x1 head: the pad channels are zero and stay zero (gamma = beta = 0)
        gamma, beta = F.pad(gamma, (0, n_pad)), F.pad(beta, (0, n_pad))
    if batch_stats:
        if sums is not None and sums.numel() == 2 * x.shape[1]:
            y, mean, rstd, _, _ = ops.bn_act_fwd_stats(x, residual, gamma, beta, sums, bn.eps, act)
        else:
            y, mean, rstd, _, _ = ops.bn_act_fwd(x, residual, gamma, beta, bn.eps, act)
        if bn.training and bn.track_running_stats:   # nn.BatchNorm2d's running-statistics update (:200), one small kernel
            with torch.no_grad():
                nbt_pending = _nbt_list()
                if nbt_pending is not None and bn.momentum is not None:
                    nbt_pending.append(bn.num_batches_tracked)
                else:
                    bn.num_batches_tracked += 1
                momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                if n_pad > 0:
                    mean, rstd = mean[:n_feat].contiguous(), rstd[:n_feat].contiguous()
                if pre_bias is not None:
                    mean = mean + pre_bias.float()
                ops.bn_update_running(bn.running_mean, bn.running_var, mean, rstd, momentum, bn.eps,
                                      x.shape[0] * x.shape[2] * x.shape[3])
        if pre_bias is not None and pre_bias.requires_grad and torch.is_grad_enabled():
            y = _AttachZeroGrad.apply(y, pre_bias)
        return y
    with torch.no_grad():
        a = gamma[:n_feat] * torch.rsqrt(bn.running_var.float() + bn.eps)
        shift = bn.running_mean.float() if pre_bias is None else bn.running_mean.float() - pre_bias.float()
        b = beta[:n_feat] - shift * a
        if n_pad > 0:
            a, b = F.pad(a, (0, n_pad)), F.pad(b, (0, n_pad))
        return ops.bn_act_apply(x, residual, a, b, act)


def _own_conv3x3(conv: nn.Conv2d, x: torch.Tensor, weight: Optional[torch.Tensor] = None) -> bool:
    """True when K7 (csrc/conv_sm100.cu) takes this convolution: the reference's conv3x3(bias=False, padding=1) on a
    bf16 channels-last activation whose geometry the tcgen05 tiling covers (all of them at 128x128 inputs except the
    3-channel stem)."""
    return (x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16
            and x.is_contiguous(memory_format=torch.channels_last)
            and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None
            and conv.weight.dtype == torch.float32
            and ops.conv3x3_shape_ok(x.shape[0], x.shape[1], conv.out_channels, x.shape[2], x.shape[3]))


def _inference_cache(module: nn.Module, slot: str, params, build):
    """Operand copies of parameters (bf16 weight tiles, the packed [3C, C] projection matrix) are rebuilt on every
    forward while training -- the optimiser changes the parameters every step.  Under ``torch.no_grad()`` (inference:
    BASELINE config 4, validation loops) they are kept on the module, keyed by every parameter's (storage address,
    version counter, shape): ``optimizer.step``, ``load_state_dict`` and ``.to()`` change the key.  Writing through
    ``param.data`` does not bump the version counter -- call ``invalidate_operand_caches(model)`` after doing that."""
    key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in params if p is not None)
    hit = module.__dict__.get(slot)
    if hit is None or hit[0] != key:
        hit = (key, build())
        module.__dict__[slot] = hit
    return hit[1]


def invalidate_operand_caches(model: nn.Module) -> None:
    """Drop the inference-mode operand copies kept on the modules of ``model`` (see _inference_cache)."""
    for m in model.modules():
        for slot in ("_mu_conv_operand", "_mu_qkv_operand"):
            m.__dict__.pop(slot, None)


def fold_skip_grads() -> bool:
    """MASKUNET_FOLD_SKIP_GRADS=0 (environment, A/B runs): leave the accumulation of the two gradients of a residual /
    skip activation to autograd (one extra pass over both) instead of adding inside the data-gradient kernels."""
    return os.environ.get("MASKUNET_FOLD_SKIP_GRADS", "1") != "0"


def conv_bn_act(conv: nn.Conv2d, bn: nn.BatchNorm2d, x: torch.Tensor, act: int,
                residual: Optional[torch.Tensor] = None, want_skip: bool = False):
    """act(BN(conv(x)) [+ residual]): one conv3x3 -> BatchNorm2d -> activation link of ade_semantic.py:198-210.
    On the production layout the convolution is our implicit-GEMM kernel and its epilogue hands the BatchNorm
    batch statistics to the fused normalise + activate kernel.

    ``want_skip``: returns (result, x_skip) -- x_skip is x for a second consumer (the residual add of :207); when the
    convolution is ours and x needs a gradient, it is the alias through which that consumer's gradient reaches the
    data-gradient kernel (ops._Conv3x3Skip), otherwise x itself."""
    x_in = x
    weight = conv.weight
    if (conv.in_channels < 8 and x.dim() == 4 and x.is_cuda and x.dtype == torch.bfloat16 and conv.bias is None
            and conv.kernel_size == (3, 3) and x.is_contiguous(memory_format=torch.channels_last)):
        # the 3-channel stem: zero-pad image and weight to 8 input channels (16-byte pixels, TMA-addressable);
        # autograd slices the weight gradient back
        # (maskunet_b200.data.to_tensor(..., pad_to=8) delivers the image already padded)
        extra = 8 - conv.in_channels
        if x.shape[1] == conv.in_channels:
            x = F.pad(x, (0, 0, 0, 0, 0, extra)).contiguous(memory_format=torch.channels_last)
        if x.shape[1] == 8:
            weight = F.pad(weight, (0, 0, 0, 0, 0, extra))
    if _own_conv3x3(conv, x, weight):
        if not torch.is_grad_enabled():
            # inference: the bf16 weight tiles are prepared once per parameter version, not once per forward
            wf = _inference_cache(conv, "_mu_conv_operand", (conv.weight,),
                                  lambda: ops.conv_prep_weights(weight.detach().contiguous(), False)[0])
            y, sums = ops.conv3x3_fwd(x, wf, bn.training)
        elif want_skip and x.requires_grad and x is x_in and fold_skip_grads():
            y, sums, x_skip = ops.conv3x3_skip(x, weight, bn.training)
            return fused_bn_act(y, bn, act, residual, sums=sums if bn.training else None), x_skip
        else:
            y, sums, _ = ops.conv3x3(x, weight, bn.training)
        out = fused_bn_act(y, bn, act, residual, sums=sums if bn.training else None)
        return (out, x_in) if want_skip else out
    out = fused_bn_act(conv(x[:, :conv.in_channels]), bn, act, residual)
    return (out, x_in) if want_skip else out


class ConvBlock(nn.Module):
    """conv3x3 -> BN -> GELU(erf) -> conv3x3 -> BN, optionally gelu(x + block(x)).  ade_semantic.py:192-210."""

    def __init__(self, in_channels, out_channels, mid_channels=None, residual=False):
        super().__init__()
        self.residual = residual
        mid = mid_channels if mid_channels else out_channels
        layers = [nn.Conv2d(in_channels, mid, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(mid), nn.GELU(),
                  nn.Conv2d(mid, out_channels, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(out_channels)]
        self.conv_block = nn.Sequential(*layers)

    def forward(self, x):
        conv1, bn1, _, conv2, bn2 = self.conv_block
        if self.residual:
            h, x_skip = conv_bn_act(conv1, bn1, x, ops.ACT_GELU, want_skip=True)
            return conv_bn_act(conv2, bn2, h, ops.ACT_GELU, residual=x_skip)
        h = conv_bn_act(conv1, bn1, x, ops.ACT_GELU)
        return conv_bn_act(conv2, bn2, h, ops.ACT_NONE)


def _dead_embedding(emb_dim, out_channels):
    # constructed by the reference (ade_semantic.py:222-225, 243-246) but never called; kept so that
    # parameter init consumes the same RNG stream and state_dict keys match
    return nn.Sequential(nn.SiLU(), nn.Linear(emb_dim, out_channels))


class DownSample(nn.Module):
    """MaxPool2d(2) -> ConvBlock(res) -> ConvBlock -> BN.  ade_semantic.py:212-229."""

    def __init__(self, in_channels, out_channels, emb_dim=256):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), ConvBlock(in_channels, in_channels, residual=True),
                                          ConvBlock(in_channels, out_channels), nn.BatchNorm2d(out_channels))
        self.emb_layer = _dead_embedding(emb_dim, out_channels)

    def forward(self, x, return_skip: bool = False):
        """``return_skip`` (our extension; the U-Net trunk passes it): returns (result, x_skip), x_skip being x for the
        skip connection -- the alias through which the UpSample's gradient of x reaches the pooling backward kernel."""
        pool, block1, block2, bn = self.maxpool_conv
        x_skip = x
        if _fast_layout(x) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            if return_skip and x.requires_grad and torch.is_grad_enabled() and fold_skip_grads():
                h, x_skip = ops.maxpool2_skip(x)
            else:
                h = ops.maxpool2(x)                  # K9, channels-last kernel
        else:
            h = pool(x)
        out = fused_bn_act(block2(block1(h)), bn, ops.ACT_NONE)
        return (out, x_skip) if return_skip else out


class UpSample(nn.Module):
    """bilinear x2 (align_corners) -> cat([skip, x]) -> ConvBlock(res) -> ConvBlock(mid=in/2) -> BN.
    ade_semantic.py:231-256."""

    def __init__(self, in_channels, out_channels, emb_dim=256):
        super().__init__()
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.out_channels = out_channels
        self.conv = nn.Sequential(ConvBlock(in_channels, in_channels, residual=True),
                                  ConvBlock(in_channels, out_channels, in_channels // 2),
                                  nn.BatchNorm2d(out_channels))
        self.emb_layer = _dead_embedding(emb_dim, out_channels)

    def forward(self, x, skip_x):
        block1, block2, bn = self.conv
        if _fast_layout(x) and _fast_layout(skip_x) and x.dtype == skip_x.dtype:
            h = ops.upsample_concat(skip_x, x)       # K10: bilinear x2 + concat in one pass
        else:
            up = self.upsample(x)
            h = torch.cat([skip_x, up.to(skip_x.dtype)], dim=1)
        return fused_bn_act(block2(block1(h)), bn, ops.ACT_NONE)


class UNet(nn.Module):
    """3-level U-Net with six Mask Attention sites.  ade_semantic.py:258-314.

    ``embed_dim`` selects the Cityscapes-instance variant (city_instance.py:216-276): two extra heads and a
    (semantic, boundary, embeddings) tuple output.  Keyword-only extras: ``compute_dtype=torch.bfloat16``
    runs activations in bf16 (fp32 master parameters, fp32 statistics) -- BASELINE.json configs 2-4;
    ``mask_mode`` is forwarded to the attention modules.
    """

    def __init__(self, c_in=3, c_out=3, embed_dim: Optional[int] = None, *,
                 compute_dtype: torch.dtype = torch.float32, mask_mode: str = "cached",
                 channels_last: bool = False):
        super().__init__()
        akw = dict(mask_mode=mask_mode)
        self.channels_last = channels_last
        self.initial_conv = ConvBlock(c_in, 64)
        self.downsample1 = DownSample(64, 128)
        self.self_attention1 = Mask2FormerAttention(128, 128, **akw)
        self.downsample2 = DownSample(128, 256)
        self.self_attention2 = Mask2FormerAttention(256, 256, **akw)
        self.downsample3 = DownSample(256, 256)
        self.self_attention3 = Mask2FormerAttention(256, 256, **akw)
        self.bottom1 = ConvBlock(256, 512)
        self.bottom2 = ConvBlock(512, 512)
        self.bottom3 = ConvBlock(512, 256)
        self.dropout = nn.Dropout(0.3)
        self.upsample1 = UpSample(512, 128)
        self.self_attention4 = Mask2FormerAttention(128, 128, **akw)
        self.upsample2 = UpSample(256, 64)
        self.self_attention5 = Mask2FormerAttention(64, 64, **akw)
        self.upsample3 = UpSample(128, 64)
        self.self_attention6 = Mask2FormerAttention(64, 64, **akw)
        self.norm = nn.LayerNorm([64, 128, 128])
        self.final_layer = nn.Sequential(nn.Conv2d(64, c_out, kernel_size=1), nn.BatchNorm2d(c_out), nn.ReLU())
        self.instance_variant = embed_dim is not None
        if self.instance_variant:  # construction order as city_instance.py:242-252
            self.boundary_head = nn.Sequential(nn.Conv2d(c_out, 32, kernel_size=3, padding=1), nn.BatchNorm2d(32),
                                               nn.ReLU(), nn.Conv2d(32, 1, kernel_size=1))
            self.embedding_head = nn.Sequential(nn.Conv2d(64, embed_dim, kernel_size=1),
                                                nn.BatchNorm2d(embed_dim), nn.ReLU())
        self.compute_dtype = compute_dtype
        self._padded_logits = None     # class-padded logits buffer of the last forward (see _conv_bn_relu)

    def attention_sites(self):
        return [getattr(self, f"self_attention{i}") for i in range(1, 7)]

    def _trunk(self, x):
        x1 = self.initial_conv(x)
        d, x1 = self.downsample1(x1, return_skip=True)      # (x1 .. x3 from here on: the skip-connection aliases)
        x2 = self.self_attention1(d)
        d, x2 = self.downsample2(x2, return_skip=True)
        x3 = self.self_attention2(d)
        d, x3 = self.downsample3(x3, return_skip=True)
        x4 = self.self_attention3(d)
        x4 = self.bottom3(self.bottom2(self.bottom1(x4)))
        h = self.self_attention4(self.dropout(self.upsample1(x4, x3)))
        h = self.self_attention5(self.dropout(self.upsample2(h, x2)))
        h = self.self_attention6(self.upsample3(h, x1))
        if _fast_layout(h) and tuple(h.shape[1:]) == tuple(self.norm.normalized_shape):
            # K11: LayerNorm([C, H, W]) on channels-last memory; parameters viewed in the same element order
            gamma = self.norm.weight.float().permute(1, 2, 0).contiguous()
            beta = self.norm.bias.float().permute(1, 2, 0).contiguous()
            return ops.sample_layernorm(h, gamma, beta, self.norm.eps)[0]
        return self.norm(h)

    @staticmethod
    def _own_conv1x1(conv, h):
        """K12: the 1x1 heads run on the tcgen05 kernels when the activation is bf16 channels-last."""
        return (h.is_cuda and h.dim() == 4 and h.dtype == torch.bfloat16
                and h.is_contiguous(memory_format=torch.channels_last)
                and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0)
                and conv.groups == 1 and conv.weight.dtype == torch.float32
                and ops.conv1x1_shape_ok(conv.in_channels, conv.out_channels, h.shape[2], h.shape[3]))

    def _conv_bn_relu(self, seq, h, keep_padded=False, min_pad=0):
        """ReLU(BN(conv(h))) of the output heads (ade_semantic.py:283-287).  On the production layout the 1x1
        convolution writes a class-padded buffer ([B, 160, H, W] for 150 classes, pad channels zero); the result is
        the first c_out channels of it, as a view."""
        conv, bn = seq[0], seq[1]
        if self._own_conv1x1(conv, h):
            n_pad = ops.pad_channels(max(conv.out_channels, min_pad))
            bias = conv.bias.float() if conv.bias is not None else None
            # (training: the epilogue of the head convolution hands the BatchNorm its batch statistics, MASKUNET_HEAD_STATS=0:
            # a separate statistics pass over the class-padded logits)
            want = bn.training and os.environ.get("MASKUNET_HEAD_STATS", "1") != "0"
            y_pad, _, sums = ops.conv1x1(h, conv.weight, bias, n_pad, want)
            out_pad = fused_bn_act(y_pad, bn, ops.ACT_RELU, sums=sums if want else None)
            if keep_padded:
                self._padded_logits = out_pad       # train.Trainer starts backward from the padded tensor
            return out_pad[:, :conv.out_channels] if n_pad > conv.out_channels else out_pad
        return fused_bn_act(conv(h), bn, ops.ACT_RELU)

    def _boundary(self, semantic):
        """boundary_head (city_instance.py:242-247, :275): conv3x3(c_out -> 32, bias) -> BN -> ReLU -> conv1x1(32 -> 1).
        On the production layout both convolutions run on our tcgen05 kernels: the semantic head left its logits in a
        64-channel class-padded buffer (pad channels exactly zero), so the 3x3 is K7's plain 64 -> 64 shape with the
        weight zero-padded (autograd slices the gradient back); its bias cancels under batch statistics and folds into
        the affine in eval mode (fused_bn_act's ``pre_bias``); the 1x1 is K12 on the 64-channel activation."""
        conv3, bn, _, conv1 = self.boundary_head
        pad = self._padded_logits
        own = (pad is not None and pad.shape[1] == 64 and pad.data_ptr() == semantic.data_ptr()
               and conv3.out_channels <= 64 and conv3.in_channels <= 64 and conv3.kernel_size == (3, 3)
               and conv3.padding == (1, 1) and conv3.weight.dtype == torch.float32
               and conv1.kernel_size == (1, 1) and conv1.weight.dtype == torch.float32
               and ops.conv3x3_shape_ok(pad.shape[0], 64, 64, pad.shape[2], pad.shape[3])
               and ops.conv1x1_shape_ok(64, conv1.out_channels, pad.shape[2], pad.shape[3]))
        if not own:
            return conv1(fused_bn_act(conv3(semantic), bn, ops.ACT_RELU))
        w3 = F.pad(conv3.weight, (0, 0, 0, 0, 0, 64 - conv3.in_channels, 0, 64 - conv3.out_channels))
        y, sums, _ = ops.conv3x3(pad, w3, bn.training)
        hid = fused_bn_act(y, bn, ops.ACT_RELU, sums=sums if bn.training else None, pre_bias=conv3.bias)
        w1 = F.pad(conv1.weight, (0, 0, 0, 0, 0, 64 - conv1.in_channels))
        n_pad = ops.pad_channels(conv1.out_channels)
        bias = conv1.bias.float() if conv1.bias is not None else None
        out = ops.conv1x1(hid, w1, bias, n_pad)[0]
        return out[:, :conv1.out_channels]

    def _heads(self, h):
        self._padded_logits = None
        if not self.instance_variant:
            return self._conv_bn_relu(self.final_layer, h, keep_padded=True)
        embeddings = self._conv_bn_relu(self.embedding_head, h)          # city_instance.py:273-276 order
        # 64-channel class padding (instead of 32 for 19 classes): the boundary head's 3x3 then is a 64 -> 64 conv
        semantic = self._conv_bn_relu(self.final_layer, h, keep_padded=True,
                                      min_pad=64 if self.final_layer[0].out_channels <= 64 else 0)
        boundary = self._boundary(semantic)
        return semantic, boundary, embeddings

    def forward(self, x):
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        outer = _nbt_list()
        _TLS.nbt = pending = [] if self.training else None
        try:
            if self.compute_dtype == torch.bfloat16:
                with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                    out = self._heads(self._trunk(x.to(torch.bfloat16)))
            else:
                out = self._heads(self._trunk(x))
            if pending:
                torch._foreach_add_(pending, 1)
        finally:
            _TLS.nbt = outer
        return out


class InstanceUNet(UNet):
    """The Cityscapes-instance model (city_instance.py:216-276): ``UNet(c_in, c_out, embed_dim=16)`` with
    boundary and embedding heads and a 3-tuple output.  Import as ``InstanceUNet as UNet`` in that script."""

    def __init__(self, c_in=3, c_out=3, embed_dim=16, **kw):
        super().__init__(c_in, c_out, embed_dim, **kw)
