"""Host-side input handling of `match()`: type checks and PIL -> normalised tensor.

Mirrors `_check_input` (`romatch/models/matcher.py:530-547`), `check_rgb` / `check_not_i16`
(`romatch/utils/utils.py:655-661`) and `get_tuple_transform_ops(resize, normalize=True)`
(`utils.py:164-173`: PIL bicubic resize -> /255 -> ImageNet mean/std).  The reference does this
on the host too; it is not part of the device hot path (SURVEY §8f rank 3 lists moving it to the
GPU as a "next" row).
"""
from __future__ import annotations

import os

import numpy as np
import torch
from PIL import Image

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def check_input(im_input):
    """Path -> RGB PIL; PIL must already be RGB; tensors must be [B,3,H,W] with H,W % 14 == 0."""
    if isinstance(im_input, (str, os.PathLike)):
        im = Image.open(im_input)
        if im.mode == "I;16":
            raise NotImplementedError("Can't handle 16 bit images")
        return im.convert("RGB")
    if isinstance(im_input, Image.Image):
        if im_input.mode != "RGB":
            raise NotImplementedError("Can't handle non-RGB images")
        return im_input
    assert isinstance(im_input, torch.Tensor), "im_input must be a string, path, or PIL image"
    B, C, H, W = im_input.shape
    assert C == 3, "im_input must be a RGB image"
    assert H % 14 == 0, "im_input must be a multiple of 14"
    assert W % 14 == 0, "im_input must be a multiple of 14"
    return im_input


def pil_to_normalized(im: Image.Image, size_hw) -> torch.Tensor:
    """Bicubic PIL resize to (h, w), scale to [0,1], ImageNet-normalise -> float32 [3,h,w]."""
    h, w = size_hw
    im = im.resize((w, h), Image.BICUBIC)           # torchvision Resize on PIL == PIL.resize
    arr = np.asarray(im, dtype=np.float32).transpose(2, 0, 1) / 255.0
    t = torch.from_numpy(np.ascontiguousarray(arr))
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32)[:, None, None]
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32)[:, None, None]
    return (t[:3] - mean) / std
