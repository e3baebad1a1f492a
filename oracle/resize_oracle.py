"""CPU restatement of the two cv2.resize calls of the reference's datasets.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): imported by tests/ and the golden generator, never by the product.

Reference call sites: /root/reference/code/ade20k/ade_semantic.py:72-73 (same in every dataset class):
    image_rgb = cv2.resize(image_rgb, (W, H), interpolation=cv2.INTER_LINEAR)      # uint8 HxWx3
    mask      = cv2.resize(mask,      (W, H), interpolation=cv2.INTER_NEAREST)     # uint8 HxW

The arithmetic lives in a third-party dependency that is not vendored: opencv-python (requirement.txt pins
opencv-python==4.10.0.84; this image has 4.13.0).  What is restated here is OpenCV's published uint8 algorithm
(modules/imgproc/src/resize.cpp):

  INTER_NEAREST   sx = min(floor(dx * (1 / (dw / sw))), sw - 1), same for rows, in double precision; no half-pixel
                  offset.
  INTER_LINEAR    half-pixel centres: f = (d + 0.5) * (s_len / d_len) - 0.5 (float32 after the double product),
                  s = floor(f), frac = f - s; columns: s < 0 -> (0, frac 0); s >= s_len - 1 -> (s_len - 1, frac 0);
                  rows: the fraction is kept and the two row indices are clipped into the image;
                  coefficients in 11-bit fixed point (round-half-even of frac * 2048, saturated to int16);
                  horizontal pass in int32 (S[s] * a0 + S[s + 1] * a1), vertical pass
                  ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2  -> uint8.
                  Exactly-2x down-scaling in both directions is routed to INTER_AREA: (a + b + c + d + 2) >> 2.

Pinned: tests/golden/resize_*.npz hold cv2's own outputs (tests/golden/make_golden.py --resize-only, generated in the
build container where cv2 is installed) and tests/test_oracle_vs_reference.py fuzzes this file against cv2 live.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _linear_taps(s_len: int, d_len: int, clamp_fraction: bool):
    """(index int64 [d_len], a0 int32, a1 int32): source tap s (and s + 1) with 11-bit fixed-point weights.
    Columns (clamp_fraction=True): a tap left of the image becomes (0, weight 1), one at or past the last column
    (s_len - 1, weight 1).  Rows (False): OpenCV keeps the fraction and only clips the two ROW INDICES into the image
    -- both taps then read the same row but are still scaled and truncated separately, which is visible in the last
    bit when the image is enlarged vertically."""
    inv = np.float64(d_len) / np.float64(s_len)
    scale = np.float64(1.0) / inv                              # scale = 1 / inv_scale, as resize() computes it
    d = np.arange(d_len, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    frac = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo = s < 0
        s[lo], frac[lo] = 0, 0.0
        hi = s >= s_len - 1
        s[hi], frac[hi] = s_len - 1, 0.0
    a0 = np.rint((np.float32(1.0) - frac) * np.float32(COEF_SCALE)).astype(np.int64)     # cvRound: half to even
    a1 = np.rint(frac * np.float32(COEF_SCALE)).astype(np.int64)
    a0 = np.clip(a0, -32768, 32767).astype(np.int32)
    a1 = np.clip(a1, -32768, 32767).astype(np.int32)
    return s, a0, a1


def resize_linear_u8(img: np.ndarray, out_hw) -> np.ndarray:
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR) for uint8 [H, W] or [H, W, C]."""
    assert img.dtype == np.uint8
    oh, ow = out_hw
    squeeze = img.ndim == 2
    src = img[:, :, None] if squeeze else img
    sh, sw, _ = src.shape
    if sh == 2 * oh and sw == 2 * ow:                          # exact 2x: OpenCV switches to the fast INTER_AREA
        s = src.astype(np.int32)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        out = out.astype(np.uint8)
        return out[:, :, 0] if squeeze else out
    xs, xa0, xa1 = _linear_taps(sw, ow, True)
    ys, ya0, ya1 = _linear_taps(sh, oh, False)
    x1 = np.minimum(xs + 1, sw - 1)
    y1 = np.clip(ys + 1, 0, sh - 1)
    ys = np.clip(ys, 0, sh - 1)
    s32 = src.astype(np.int32)
    # horizontal pass on the two source rows of every output row
    def hrow(rows):
        r = s32[rows]                                          # [oh, sw, C]
        return r[:, xs] * xa0[None, :, None] + r[:, x1] * xa1[None, :, None]
    r0, r1 = hrow(ys), hrow(y1)
    b0, b1 = ya0[:, None, None], ya1[:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def resize_nearest_u8(img: np.ndarray, out_hw) -> np.ndarray:
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_NEAREST) for uint8 [H, W] or [H, W, C]."""
    oh, ow = out_hw
    sh, sw = img.shape[:2]
    ifx = np.float64(1.0) / (np.float64(ow) / np.float64(sw))
    ify = np.float64(1.0) / (np.float64(oh) / np.float64(sh))
    xs = np.minimum(np.floor(np.arange(ow, dtype=np.float64) * ifx).astype(np.int64), sw - 1)
    ys = np.minimum(np.floor(np.arange(oh, dtype=np.float64) * ify).astype(np.int64), sh - 1)
    return img[ys][:, xs]
