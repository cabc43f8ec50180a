"""ORACLE (test infrastructure only -- the product path never imports this package).

CPU restatement of the bilinear image resize on the input side of the path. The reference resizes with
detectron2's ResizeShortestEdge / ResizeTransform (tools/plain_train_net.py:293-298, dafne/modeling/tta.py:76-93), which
for uint8 HWC images is `PIL.Image.resize((w, h), Image.BILINEAR)`. Pillow is a third-party dependency of the reference
(not in its tree; this image ships Pillow 12.2.0); its published algorithm (src/libImaging/Resample.c, 8-bit path) is:

  per output index xx of an axis: center = (xx + 0.5) * scale, scale = in / out, support = max(scale, 1) (bilinear),
  xmin = max(0, int(center - support + 0.5)), xmax = min(in, int(center + support + 0.5)) - xmin,
  w_x = triangle((x + xmin - center + 0.5) / max(scale, 1)), normalised by their sum (all in double),
  fixed point kk = int(0.5 + w * 2^22); out = clip8((2^21 + sum in[x + xmin] * kk[x]) >> 22);
  horizontal pass first (into an 8-bit intermediate), then vertical; a pass whose size does not change is skipped.

Pinned against Pillow itself (tests/test_resize.py::test_oracle_matches_pillow), bit for bit.
"""
from __future__ import annotations

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _coefficients(in_size: int, out_size: int):
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmins = np.zeros(out_size, np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = np.zeros(ksize)
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            w = 1.0 - abs(v) if abs(v) < 1.0 else 0.0
            k[x] = w
            ww += w
        if ww != 0.0:
            k[:xmax] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        xmins[xx] = xmin
    return xmins, kk


def _resample_last_axis(planes: np.ndarray, out_size: int) -> np.ndarray:
    """planes: [..., in_size] uint8 -> [..., out_size] uint8."""
    in_size = planes.shape[-1]
    xmins, kk = _coefficients(in_size, out_size)
    out = np.empty(planes.shape[:-1] + (out_size,), np.uint8)
    src = planes.astype(np.int64)
    for xx in range(out_size):
        acc = np.full(planes.shape[:-1], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(kk.shape[1]):
            if kk[xx, x] != 0:
                acc += src[..., xmins[xx] + x] * kk[xx, x]
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(image_chw: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """image_chw: [C, H, W] uint8 -> [C, new_h, new_w] uint8, like PIL.Image.resize((new_w, new_h), BILINEAR)."""
    assert image_chw.dtype == np.uint8 and image_chw.ndim == 3
    out = image_chw
    if new_w != out.shape[2]:
        out = _resample_last_axis(out, new_w)
    if new_h != out.shape[1]:
        out = np.ascontiguousarray(_resample_last_axis(np.ascontiguousarray(out.transpose(0, 2, 1)), new_h).transpose(0, 2, 1))
    return out
