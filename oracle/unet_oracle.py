"""CPU restatement of the reference U-Net around the Mask Attention Module.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``/root/reference/code/ade20k/ade_semantic.py:192-314`` (ConvBlock,
DownSample, UpSample, UNet) and the three-output variant at
``/root/reference/code/cityscapes/city_instance.py:216-276``.  The network is
restated *functionally* over a flat ``{state_dict key: tensor}`` mapping so the
oracle, the reference and the CUDA modules can all exchange one state dict.

``init_state`` draws the initial parameters by constructing the same torch
layers in the reference's construction order, so after the same
``torch.manual_seed`` it yields the reference's initial state_dict (checked
against the reference in tests/test_oracle_vs_reference.py and against
tests/golden/unet_semantic.npz).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mask_attention_oracle as mao

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
ATTN_SITES = (  # name, channels, token-grid side for a 128x128 input (ade_semantic.py:263-280)
    ("self_attention1", 128, 64), ("self_attention2", 256, 32), ("self_attention3", 256, 16),
    ("self_attention4", 128, 32), ("self_attention5", 64, 64), ("self_attention6", 64, 128),
)


# ------------------------------------------------------------------ construction recipe
def _conv_block_recipe(prefix, cin, cout, mid=None):
    mid = mid or cout
    return [(f"{prefix}.conv_block.0", "conv3", (cin, mid)), (f"{prefix}.conv_block.1", "bn", (mid,)),
            (f"{prefix}.conv_block.3", "conv3", (mid, cout)), (f"{prefix}.conv_block.4", "bn", (cout,))]


def _down_recipe(prefix, cin, cout):
    return (_conv_block_recipe(f"{prefix}.maxpool_conv.1", cin, cin)
            + _conv_block_recipe(f"{prefix}.maxpool_conv.2", cin, cout)
            + [(f"{prefix}.maxpool_conv.3", "bn", (cout,)), (f"{prefix}.emb_layer.1", "linear", (256, cout))])


def _up_recipe(prefix, cin, cout):
    return (_conv_block_recipe(f"{prefix}.conv.0", cin, cin)
            + _conv_block_recipe(f"{prefix}.conv.1", cin, cout, cin // 2)
            + [(f"{prefix}.conv.2", "bn", (cout,)), (f"{prefix}.emb_layer.1", "linear", (256, cout))])


def _attn_recipe(prefix, c):
    return [(f"{prefix}.query", "linear", (c, c)), (f"{prefix}.key", "linear", (c, c)),
            (f"{prefix}.value", "linear", (c, c)), (f"{prefix}.norm", "ln", ((c,),))]


def unet_recipe(c_in: int, c_out: int, variant: str = "semantic", embed_dim: int = 16):
    """Layers in the reference's construction order (ade_semantic.py:260-287)."""
    r = []
    r += _conv_block_recipe("initial_conv", c_in, 64)
    r += _down_recipe("downsample1", 64, 128) + _attn_recipe("self_attention1", 128)
    r += _down_recipe("downsample2", 128, 256) + _attn_recipe("self_attention2", 256)
    r += _down_recipe("downsample3", 256, 256) + _attn_recipe("self_attention3", 256)
    r += _conv_block_recipe("bottom1", 256, 512) + _conv_block_recipe("bottom2", 512, 512)
    r += _conv_block_recipe("bottom3", 512, 256)
    r += _up_recipe("upsample1", 512, 128) + _attn_recipe("self_attention4", 128)
    r += _up_recipe("upsample2", 256, 64) + _attn_recipe("self_attention5", 64)
    r += _up_recipe("upsample3", 128, 64) + _attn_recipe("self_attention6", 64)
    r += [("norm", "ln", ((64, 128, 128),))]
    r += [("final_layer.0", "conv1b", (64, c_out)), ("final_layer.1", "bn", (c_out,))]
    if variant == "instance":  # city_instance.py:242-252
        r += [("boundary_head.0", "conv3b", (c_out, 32)), ("boundary_head.1", "bn", (32,)),
              ("boundary_head.3", "conv1b", (32, 1))]
        r += [("embedding_head.0", "conv1b", (64, embed_dim)), ("embedding_head.1", "bn", (embed_dim,))]
    return r


def init_state(c_in: int = 3, c_out: int = 3, variant: str = "semantic", embed_dim: int = 16
               ) -> "OrderedDict[str, torch.Tensor]":
    """Initial state_dict drawn with torch's default initialisers in reference order."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, kind, args in unet_recipe(c_in, c_out, variant, embed_dim):
        if kind == "conv3":
            layer = nn.Conv2d(args[0], args[1], kernel_size=3, padding=1, bias=False)
        elif kind == "conv3b":
            layer = nn.Conv2d(args[0], args[1], kernel_size=3, padding=1)
        elif kind == "conv1b":
            layer = nn.Conv2d(args[0], args[1], kernel_size=1)
        elif kind == "bn":
            layer = nn.BatchNorm2d(args[0])
        elif kind == "linear":
            layer = nn.Linear(args[0], args[1])
        elif kind == "ln":
            layer = nn.LayerNorm(list(args[0]))
        else:  # pragma: no cover
            raise ValueError(kind)
        for k, v in layer.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


def trainable_keys(sd) -> list:
    return [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]


# ------------------------------------------------------------------ functional forward
def _bn(x, sd, name, training, update_stats):
    rm, rv = sd[f"{name}.running_mean"], sd[f"{name}.running_var"]
    if training and not update_stats:
        rm, rv = rm.clone(), rv.clone()
    return F.batch_norm(x, rm, rv, sd[f"{name}.weight"], sd[f"{name}.bias"], training, BN_MOMENTUM, BN_EPS)


def conv_block(x, sd, prefix, residual=False, training=False, update_stats=False):
    """ade_semantic.py:192-210: conv3x3 -> BN -> GELU(erf) -> conv3x3 -> BN [, gelu(x + .)]."""
    h = F.conv2d(x, sd[f"{prefix}.conv_block.0.weight"], None, padding=1)
    h = _bn(h, sd, f"{prefix}.conv_block.1", training, update_stats)
    h = F.gelu(h)
    h = F.conv2d(h, sd[f"{prefix}.conv_block.3.weight"], None, padding=1)
    h = _bn(h, sd, f"{prefix}.conv_block.4", training, update_stats)
    return F.gelu(x + h) if residual else h


def down_sample(x, sd, prefix, training=False, update_stats=False):
    """ade_semantic.py:212-229 (emb_layer is constructed but never used)."""
    h = F.max_pool2d(x, 2)
    h = conv_block(h, sd, f"{prefix}.maxpool_conv.1", True, training, update_stats)
    h = conv_block(h, sd, f"{prefix}.maxpool_conv.2", False, training, update_stats)
    return _bn(h, sd, f"{prefix}.maxpool_conv.3", training, update_stats)


def up_sample(x, skip, sd, prefix, training=False, update_stats=False):
    """ade_semantic.py:231-256: bilinear x2 (align_corners=True) -> cat([skip, x]) -> convs -> BN."""
    h = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    h = torch.cat([skip, h], dim=1)
    h = conv_block(h, sd, f"{prefix}.conv.0", True, training, update_stats)
    h = conv_block(h, sd, f"{prefix}.conv.1", False, training, update_stats)
    return _bn(h, sd, f"{prefix}.conv.2", training, update_stats)


def mask_attention(x, sd, prefix, keep):
    """ade_semantic.py:163-190 through the explicit restatement."""
    B, C, H, W = x.shape
    params = {k: sd[f"{prefix}.{k}"] for k in ("query.weight", "query.bias", "key.weight", "key.bias",
                                               "value.weight", "value.bias", "norm.weight", "norm.bias")}
    out = mao.attention_forward(x.contiguous(), params, keep, dtype=x.dtype)
    return mao.module_output(out["y"], C, H, W)


def mask_attention_reference_ops(x, sd, prefix, keep):
    """ade_semantic.py:163-190 in the reference's OWN op sequence -- ``nn.Linear`` (F.linear), ``matmul``, true
    division by ``C ** 0.5``, ``+`` the cached expanded 0 / -inf mask, ``F.softmax``, ``matmul``, residual,
    ``nn.LayerNorm`` (F.layer_norm), ``.view`` -- under autograd.  This is what the CPU baseline times: the explicit
    restatement above (max / exp / sum / div as separate N x N tensors, hand-derived backward) is an independent
    statement of the maths for parity and costs ~1.7x the reference's own sequence on the same cores."""
    B, C, H, W = x.shape
    N = H * W
    xt = x.view(B, C, N).permute(0, 2, 1)                                           # :168
    q = F.linear(xt, sd[f"{prefix}.query.weight"], sd[f"{prefix}.query.bias"])      # :170
    k = F.linear(xt, sd[f"{prefix}.key.weight"], sd[f"{prefix}.key.bias"])          # :171
    v = F.linear(xt, sd[f"{prefix}.value.weight"], sd[f"{prefix}.value.bias"])      # :172
    scores = torch.matmul(q, k.transpose(-2, -1))                                   # :174
    scores = scores / (C ** 0.5)                                                    # :175
    mask = mao.expand_bias(mao.additive_bias(keep), N)                              # :179-181 (cached by the module)
    scores = scores + mask                                                          # :183
    w = F.softmax(scores, dim=-1)                                                   # :185
    out = torch.matmul(w, v)                                                        # :186
    out = out + xt                                                                  # :187
    out = F.layer_norm(out, [C], sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"], 1e-5)   # :188
    return out.view(B, C, H, W)                                                     # :190


def draw_keeps(batch: int, image_hw: Tuple[int, int] = (128, 128)) -> Dict[str, torch.Tensor]:
    """Per-site keep masks drawn in forward order with the reference's randint call (:178)."""
    keeps = {}
    scale = image_hw[0] / 128.0
    order = ("self_attention1", "self_attention2", "self_attention3",
             "self_attention4", "self_attention5", "self_attention6")
    sides = dict((n, s) for n, _, s in ATTN_SITES)
    for name in order:
        side_h = int(sides[name] * scale)
        side_w = int(sides[name] * image_hw[1] / 128.0)
        keeps[name] = mao.binarize_mask(mao.draw_mask_bits(batch, side_h, side_w))
    return keeps


class LazyKeeps(dict):
    """Keep masks drawn on first use, i.e. in forward order between the dropout draws -- exactly where the
    reference's modules call ``torch.randint`` (ade_semantic.py:177-178) -- and cached afterwards."""

    def __init__(self, batch: int, image_hw: Tuple[int, int] = (128, 128)):
        super().__init__()
        self.batch, self.image_hw = batch, image_hw

    def __missing__(self, name):
        side = dict((n, s) for n, _, s in ATTN_SITES)[name]
        h, w = int(side * self.image_hw[0] / 128.0), int(side * self.image_hw[1] / 128.0)
        self[name] = mao.binarize_mask(mao.draw_mask_bits(self.batch, h, w))
        return self[name]


def unet_forward(sd, x, keeps: Dict[str, torch.Tensor], training: bool = False,
                 dropout_p: float = 0.0, update_stats: bool = False, variant: str = "semantic",
                 attention=None):
    """ade_semantic.py:289-314 (semantic) / city_instance.py:253-276 (instance, 3 outputs).

    ``dropout_p`` defaults to 0 (parity runs); pass 0.3 with ``training=True`` for the
    timed CPU baseline, where it consumes the torch RNG as nn.Dropout does (:273,304,307).
    """
    t, u = training, update_stats
    mask_attention = attention or globals()["mask_attention"]      # explicit restatement unless told otherwise
    x1 = conv_block(x, sd, "initial_conv", False, t, u)
    x2 = mask_attention(down_sample(x1, sd, "downsample1", t, u), sd, "self_attention1", keeps["self_attention1"])
    x3 = mask_attention(down_sample(x2, sd, "downsample2", t, u), sd, "self_attention2", keeps["self_attention2"])
    x4 = mask_attention(down_sample(x3, sd, "downsample3", t, u), sd, "self_attention3", keeps["self_attention3"])
    x4 = conv_block(x4, sd, "bottom1", False, t, u)
    x4 = conv_block(x4, sd, "bottom2", False, t, u)
    x4 = conv_block(x4, sd, "bottom3", False, t, u)
    h = up_sample(x4, x3, sd, "upsample1", t, u)
    h = F.dropout(h, dropout_p, training)
    h = mask_attention(h, sd, "self_attention4", keeps["self_attention4"])
    h = up_sample(h, x2, sd, "upsample2", t, u)
    h = F.dropout(h, dropout_p, training)
    h = mask_attention(h, sd, "self_attention5", keeps["self_attention5"])
    h = up_sample(h, x1, sd, "upsample3", t, u)
    h = mask_attention(h, sd, "self_attention6", keeps["self_attention6"])
    h = F.layer_norm(h, list(sd["norm.weight"].shape), sd["norm.weight"], sd["norm.bias"], 1e-5)

    def head(prefix, inp):  # conv1x1(bias) -> BN -> ReLU   (:283-287)
        o = F.conv2d(inp, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"])
        return F.relu(_bn(o, sd, f"{prefix}.1", t, u))

    if variant == "instance":
        emb = head("embedding_head", h)
        sem = head("final_layer", h)
        b = F.conv2d(sem, sd["boundary_head.0.weight"], sd["boundary_head.0.bias"], padding=1)
        b = F.relu(_bn(b, sd, "boundary_head.1", t, u))
        b = F.conv2d(b, sd["boundary_head.3.weight"], sd["boundary_head.3.bias"])
        return sem, b, emb
    return head("final_layer", h)


# ------------------------------------------------------------------ CPU baseline train step
class OracleTrainer:
    """The reference's timed loop body (ade_semantic.py:394-401) on the restated network:
    zero_grad -> forward -> CrossEntropyLoss -> backward -> AdamW(lr 5e-5, wd 1e-1)."""

    def __init__(self, c_in=3, c_out=150, lr=5e-5, weight_decay=1e-1, seed=42, dropout_p=0.3,
                 reference_ops: bool = False):
        """reference_ops=True: the attention sites run the reference's own op sequence
        (mask_attention_reference_ops) -- the CPU baseline that bench.py times; False: the explicit restatement."""
        self.attention = mask_attention_reference_ops if reference_ops else None
        torch.manual_seed(seed)
        self.sd = init_state(c_in, c_out)
        self.params = [self.sd[k].requires_grad_(True) for k in trainable_keys(self.sd)]
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)
        self.keeps: Optional[Dict[str, torch.Tensor]] = None
        self.dropout_p = dropout_p

    def step(self, images: torch.Tensor, labels: torch.Tensor) -> float:
        if self.keeps is None:  # cached on first forward, as self.mask is (:177); drawn in forward order
            self.keeps = LazyKeeps(images.shape[0], tuple(images.shape[-2:]))
        self.opt.zero_grad(set_to_none=True)
        logits = unet_forward(self.sd, images, self.keeps, training=True,
                              dropout_p=self.dropout_p, update_stats=True, attention=self.attention)
        loss = F.cross_entropy(logits, labels)
        loss.backward()
        self.opt.step()
        return float(loss.detach())
