"""Seeded synthetic weights in the reference's state-dict layout.

The reference fetches `roma_outdoor.pth` / `dinov2_vitl14_pretrain.pth` from URLs
(`romatch/models/model_zoo/__init__.py:6-15,42-49`); neither this container nor the GPU box has
a network, so every parity test, golden vector and benchmark in this repository runs on weights
produced here.  The generator is pure torch-CPU with an explicit `torch.Generator`, so the same
seed gives bit-identical tensors on every machine with the same torch build, and the resulting
dicts load into the *unmodified* reference with `strict=True` (that is how `tests/golden` was
produced).

All BatchNorm running statistics, LayerScale gammas and biases are randomised so that BN folding,
LayerScale and bias epilogues are actually exercised by the parity tests (defaults of 0/1 would hide
bugs there).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import arch


def _fan_in(shape):
    n = 1
    for d in shape[1:]:
        n *= d
    return max(n, 1)


def _fill(kind: str, shape, g: torch.Generator) -> torch.Tensor:
    def randn(std):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    def uniform(lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    if kind in ("conv", "linear", "pos_conv", "disp_emb"):
        return randn(1.0 / math.sqrt(_fan_in(shape)))
    if kind in ("conv_relu", "dwconv"):
        return randn(math.sqrt(2.0 / _fan_in(shape)))
    if kind == "to_out":                     # peaky (realistic) anchor distribution: logit std ~4
        return randn(4.0 / math.sqrt(_fan_in(shape)))
    if kind == "proj":                       # keeps projected features (and local correlations) O(1)
        return randn(0.35 / math.sqrt(_fan_in(shape)))
    if kind == "out_conv":                   # keeps per-scale flow/certainty updates moderate
        return randn(0.3 / math.sqrt(_fan_in(shape)))
    if kind == "bias":
        return randn(0.05)
    if kind in ("bn_w", "bn_var", "ln_w"):
        return uniform(0.8, 1.2)
    if kind in ("bn_b", "bn_mean", "ln_b"):
        return randn(0.1)
    if kind == "bn_count":
        return torch.zeros((), dtype=torch.int64)
    if kind == "ls":
        return uniform(0.5, 1.0)
    if kind == "token":
        return randn(0.02)
    if kind == "zeros":
        return torch.zeros(shape, dtype=torch.float32)
    raise KeyError(kind)


def make_matcher_weights(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the 603 tensors of `RegressionMatcher.state_dict()`."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * (seed + 1))
    sd = OrderedDict((k, _fill(kind, shape, g)) for k, shape, kind in arch.matcher_param_specs())
    # the certainty logit shares `to_out` with the 4096 anchor logits; keep it O(1) so that the final
    # sigmoid is not saturated and certainty errors stay visible in parity tests
    sd["decoder.embedding_decoder.to_out.weight"][-1] *= 0.125
    for s in arch.SCALES:                    # centre the accumulated certainty logit near 0
        sd[f"decoder.conv_refiner.{s}.out_conv.bias"][2] += 0.6
    return sd


def make_dinov2_weights(seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """State dict of DINOv2 ViT-L/14 as `vit_large(...).state_dict()` lays it out."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * (seed + 1) + 17)
    return OrderedDict((k, _fill(kind, shape, g)) for k, shape, kind in arch.dinov2_param_specs())


def make_weights(seed: int = 0):
    return make_matcher_weights(seed), make_dinov2_weights(seed)


def make_pair(batch: int, coarse, upsample, seed: int = 1):
    """Synthetic N(0,1) image tensors, the distribution the reference's own timing tests use
    (`tests/test_roma_upsample_inference_time.py:9-12`): (A, B, A_high, B_high).  Resolutions are ints (square) or (h, w)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    ch, cw = (coarse, coarse) if isinstance(coarse, int) else coarse
    a = torch.randn(batch, 3, ch, cw, generator=g)
    b = torch.randn(batch, 3, ch, cw, generator=g)
    if upsample is None:
        return a, b, None, None
    uh, uw = (upsample, upsample) if isinstance(upsample, int) else upsample
    ah = torch.randn(batch, 3, uh, uw, generator=g)
    bh = torch.randn(batch, 3, uh, uw, generator=g)
    return a, b, ah, bh


def make_pil_pair(seed: int = 3, size_a=(200, 150), size_b=(180, 220)):
    """Two seeded RGB PIL images of different sizes (width, height) for the PIL/path input route
    (`matcher.py:806-816`): smooth random blobs so that the bicubic resize is well exercised."""
    import numpy as np
    from PIL import Image

    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out = []
    for (w, h) in (size_a, size_b):
        low = torch.rand(1, 3, 12, 16, generator=g)
        img = torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=False)
        arr = (img[0].clamp(0, 1) * 255).round().to(torch.uint8).permute(1, 2, 0).numpy()
        out.append(Image.fromarray(np.ascontiguousarray(arr), "RGB"))
    return out[0], out[1]
