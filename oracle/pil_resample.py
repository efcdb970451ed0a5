"""CPU restatement of the image preprocessing in front of `match()` -- TEST INFRASTRUCTURE ONLY.

The reference turns a PIL image into the network input with `get_tuple_transform_ops(resize=(h, w),
normalize=True)` (`romatch/utils/utils.py:164-173`, called at `romatch/models/matcher.py:812-815,855-866`):
`torchvision.transforms.Resize(size, BICUBIC)` on a PIL image, which is `PIL.Image.resize((w, h), BICUBIC)`
(`utils.py:233-238`), then `np.array(im, float32) / 255` (`utils.py:175-183`) and ImageNet mean/std
(`utils.py:250-260`).

The resize itself lives in a third-party dependency that is not under /root/reference: Pillow (12.2.0 in this image),
`src/libImaging/Resample.c`.  Its published algorithm for 8-bit images is restated here:
  * per output coordinate a window [xmin, xmin+n) of input samples, n <= ksize = 2*ceil(support)+1 with
    support = 2 * max(1, in/out) for the bicubic filter (a = -0.5), weights normalised to sum 1 in double precision
    (`precompute_coeffs`);
  * weights rounded to 22-bit fixed point, half away from zero (`normalize_coeffs_8bpc`, PRECISION_BITS = 32-8-2);
  * horizontal pass first, then vertical, each `clip8((2^21 + sum(pixel * k)) >> 22)` with a uint8 image in between
    (`ImagingResampleHorizontal_8bpc` / `ImagingResampleVertical_8bpc`); a pass whose size does not change is skipped.
Parity is pinned against Pillow itself (installed here and on the GPU box): tests/test_preprocess.py compares this
restatement with `PIL.Image.resize` bit for bit, and the CUDA path with both.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coeffs(in_size: int, out_size: int):
    """`precompute_coeffs` + `normalize_coeffs_8bpc` for the full-image box: (ksize, bounds[out,2], kk[out,ksize] int32)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8-bit resampling pass along axis 1 of a [rows, n, C] uint8 image."""
    n = img.shape[1]
    ksize, bounds, kk = coeffs(n, out_size)
    out = np.empty((img.shape[0], out_size, img.shape[2]), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, cnt = bounds[xx]
        acc = (src[:, xmin:xmin + cnt, :] * kk[xx, :cnt].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, size_hw) -> np.ndarray:
    """`PIL.Image.resize((w, h), BICUBIC)` of an RGB uint8 [H, W, 3] array."""
    h, w = size_hw
    out = img
    if w != img.shape[1]:
        out = _pass(out, w)
    if h != img.shape[0]:
        out = _pass(out.transpose(1, 0, 2), h).transpose(1, 0, 2)
    return np.ascontiguousarray(out)


def preprocess(img: np.ndarray, size_hw) -> np.ndarray:
    """uint8 [H, W, 3] -> float32 [3, h, w], the tensor `get_tuple_transform_ops(resize, normalize=True)` returns."""
    x = resize_bicubic_u8(img, size_hw).astype(np.float32).transpose(2, 0, 1)
    x /= np.float32(255.0)
    mean = np.asarray(IMAGENET_MEAN, dtype=np.float32)[:, None, None]
    std = np.asarray(IMAGENET_STD, dtype=np.float32)[:, None, None]
    return (x - mean) / std
