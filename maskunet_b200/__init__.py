"""maskunet_b200 -- B200-native Mask Attention Module and the U-Net blocks around it.

Importing this package loads libmaskunet_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/maskunet_b200.h).  There is no CPU or eager-PyTorch fallback for the attention path: a missing
library raises ImportError, a CPU tensor raises RuntimeError.
"""
from . import _lib

_lib.load()  # fail loudly, now, if the CUDA library is absent

from . import ops  # noqa: E402  (registers the maskunet:: operators)
from .modules import (ConvBlock, DownSample, InstanceUNet, Mask2FormerAttention, UNet,  # noqa: E402
                      UpSample)

from . import checkpoint  # noqa: E402
from . import data  # noqa: E402  (device ToTensor + prefetcher, ade_semantic.py:56-97)
from . import query_attention  # noqa: E402  (generalised mode of the kernel sweep)
from .losses import InstanceContrastiveLoss  # noqa: E402  (device version of coco_panoptic.py:482-521)
from .ops import mean_iou, segmentation_argmax  # noqa: E402  (device versions of ade_semantic.py:128-146)
from ._lib import is_deterministic, set_deterministic  # noqa: E402  (fixed-order accumulation across CTAs)

__all__ = ["Mask2FormerAttention", "ConvBlock", "DownSample", "UpSample", "UNet", "InstanceUNet", "ops",
           "mean_iou", "segmentation_argmax", "checkpoint", "InstanceContrastiveLoss", "data", "query_attention",
           "set_deterministic", "is_deterministic"]
__version__ = "0.1.0"
