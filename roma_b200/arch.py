"""Architecture constants of the RoMa dense-matching path and its state-dict layout.

Nothing here is computed from the reference at run time: the numbers restate what
`romatch/models/model_zoo/roma_models.py:71-181` (decoder, refiners, GP, proj),
`romatch/models/encoders.py:13,35-42` (VGG19-BN features[:40], DINOv2 ViT-L/14) and
`romatch/models/transformer/dinov2.py:333` (vit_large) wire up, so that weights in the
reference's key layout can be loaded, generated and packed on a box that has no copy of
the reference.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, Tuple

# ---- VGG19-BN features[:40] (encoders.py:13): conv indices, (cin, cout); pools at 6,13,26,39
VGG_CONVS: Tuple[Tuple[int, int, int], ...] = (
    (0, 3, 64), (3, 64, 64),
    (7, 64, 128), (10, 128, 128),
    (14, 128, 256), (17, 256, 256), (20, 256, 256), (23, 256, 256),
    (27, 256, 512), (30, 512, 512), (33, 512, 512), (36, 512, 512),
)
VGG_POOLS = (6, 13, 26, 39)          # taps are the *inputs* of these pools (encoders.py:22-26)
VGG_TAP_CHANNELS = {1: 64, 2: 128, 4: 256, 8: 512}

# ---- DINOv2 ViT-L/14 (encoders.py:35-42, dinov2.py:333-346)
VIT_DIM = 1024
VIT_DEPTH = 24
VIT_HEADS = 16
VIT_MLP = 4096
VIT_PATCH = 14
VIT_POS_GRID = 37                      # img_size 518 / 14
VIT_LN_EPS = 1e-6                      # dinov2.py:88

# ---- coarse matcher (roma_models.py:71-84, 140-155)
GP_DIM = 512
FEAT_DIM = 512
DEC_DIM = GP_DIM + FEAT_DIM            # 1024
DEC_DEPTH = 5
DEC_HEADS = 8
DEC_MLP = 4096
DEC_LN_EPS = 1e-5                      # nn.LayerNorm default (block.py:50)
CLS_RES = 64
CLS_OUT = CLS_RES * CLS_RES + 1        # 4097
GP_TEMPERATURE = 0.2
GP_SIGMA_NOISE = 0.1                   # matcher.py:214
GP_COS_EPS = 1e-6                      # matcher.py:191
BN_EPS = 1e-5

SCALES = (16, 8, 4, 2, 1)
UPSAMPLE_SCALES = (8, 4, 2, 1)         # matcher.py:407
REFINE_INIT = 4                        # matcher.py:359


@dataclass(frozen=True)
class RefinerSpec:
    scale: int
    feat: int          # channels of the projected feature (x and x_hat)
    emb: int           # displacement-embedding channels
    radius: int        # local-correlation radius (0 = none)

    @property
    def k(self) -> int:
        return (2 * self.radius + 1) ** 2 if self.radius else 0

    @property
    def channels(self) -> int:
        return 2 * self.feat + self.emb + self.k


# roma_models.py:103-139
REFINERS = {
    16: RefinerSpec(16, 512, 128, 7),   # 1377
    8: RefinerSpec(8, 512, 64, 3),      # 1137
    4: RefinerSpec(4, 256, 32, 2),      # 569
    2: RefinerSpec(2, 64, 16, 0),       # 144
    1: RefinerSpec(1, 9, 6, 0),         # 24
}
REFINER_HIDDEN_BLOCKS = 8

# roma_models.py:156-169 : proj[s] = Conv2d(cin, cout, 1) + BatchNorm2d(cout)
PROJ = {16: (1024, 512), 8: (512, 512), 4: (256, 256), 2: (128, 64), 1: (64, 9)}


def matcher_param_specs() -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) in the order of the reference's `RegressionMatcher.state_dict()`.

    kind in {conv, dwconv, linear, bias, bn_w, bn_b, bn_mean, bn_var, bn_count, ln_w, ln_b,
    to_out, out_conv, pos_conv, disp_emb}
    """
    for idx, cin, cout in VGG_CONVS:
        p = f"encoder.cnn.layers.{idx}"
        yield f"{p}.weight", (cout, cin, 3, 3), "conv_relu"
        yield f"{p}.bias", (cout,), "bias"
        yield from _bn(f"encoder.cnn.layers.{idx + 1}", cout)
    for i in range(DEC_DEPTH):
        p = f"decoder.embedding_decoder.blocks.{i}"
        yield f"{p}.norm1.weight", (DEC_DIM,), "ln_w"
        yield f"{p}.norm1.bias", (DEC_DIM,), "ln_b"
        yield f"{p}.attn.qkv.weight", (3 * DEC_DIM, DEC_DIM), "linear"
        yield f"{p}.attn.proj.weight", (DEC_DIM, DEC_DIM), "linear"
        yield f"{p}.attn.proj.bias", (DEC_DIM,), "bias"
        yield f"{p}.norm2.weight", (DEC_DIM,), "ln_w"
        yield f"{p}.norm2.bias", (DEC_DIM,), "ln_b"
        yield f"{p}.mlp.fc1.weight", (DEC_MLP, DEC_DIM), "linear"
        yield f"{p}.mlp.fc1.bias", (DEC_MLP,), "bias"
        yield f"{p}.mlp.fc2.weight", (DEC_DIM, DEC_MLP), "linear"
        yield f"{p}.mlp.fc2.bias", (DEC_DIM,), "bias"
    yield "decoder.embedding_decoder.to_out.weight", (CLS_OUT, DEC_DIM), "to_out"
    yield "decoder.embedding_decoder.to_out.bias", (CLS_OUT,), "bias"
    yield "decoder.gps.16.pos_conv.weight", (GP_DIM, 2, 1, 1), "pos_conv"
    yield "decoder.gps.16.pos_conv.bias", (GP_DIM,), "bias"
    for s in SCALES:
        cin, cout = PROJ[s]
        yield f"decoder.proj.{s}.0.weight", (cout, cin, 1, 1), "proj"
        yield f"decoder.proj.{s}.0.bias", (cout,), "bias"
        yield from _bn(f"decoder.proj.{s}.1", cout)
    for s in SCALES:
        spec = REFINERS[s]
        c = spec.channels
        blocks = ["block1"] + [f"hidden_blocks.{j}" for j in range(REFINER_HIDDEN_BLOCKS)]
        for blk in blocks:
            p = f"decoder.conv_refiner.{s}.{blk}"
            yield f"{p}.0.weight", (c, 1, 5, 5), "dwconv"
            yield f"{p}.0.bias", (c,), "bias"
            yield from _bn(f"{p}.1", c)
            yield f"{p}.3.weight", (c, c, 1, 1), "conv"
            yield f"{p}.3.bias", (c,), "bias"
        yield f"decoder.conv_refiner.{s}.out_conv.weight", (3, c, 1, 1), "out_conv"
        yield f"decoder.conv_refiner.{s}.out_conv.bias", (3,), "bias"
        yield f"decoder.conv_refiner.{s}.disp_emb.weight", (spec.emb, 2, 1, 1), "disp_emb"
        yield f"decoder.conv_refiner.{s}.disp_emb.bias", (spec.emb,), "bias"


def _bn(prefix: str, c: int):
    yield f"{prefix}.weight", (c,), "bn_w"
    yield f"{prefix}.bias", (c,), "bn_b"
    yield f"{prefix}.running_mean", (c,), "bn_mean"
    yield f"{prefix}.running_var", (c,), "bn_var"
    yield f"{prefix}.num_batches_tracked", (), "bn_count"


def dinov2_param_specs() -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) in the order of `vit_large(...).state_dict()` (dinov2.py:43-155)."""
    yield "cls_token", (1, 1, VIT_DIM), "token"
    yield "pos_embed", (1, 1 + VIT_POS_GRID * VIT_POS_GRID, VIT_DIM), "token"
    yield "mask_token", (1, VIT_DIM), "zeros"
    yield "patch_embed.proj.weight", (VIT_DIM, 3, VIT_PATCH, VIT_PATCH), "conv"
    yield "patch_embed.proj.bias", (VIT_DIM,), "bias"
    for i in range(VIT_DEPTH):
        p = f"blocks.{i}"
        yield f"{p}.norm1.weight", (VIT_DIM,), "ln_w"
        yield f"{p}.norm1.bias", (VIT_DIM,), "ln_b"
        yield f"{p}.attn.qkv.weight", (3 * VIT_DIM, VIT_DIM), "linear"
        yield f"{p}.attn.qkv.bias", (3 * VIT_DIM,), "bias"
        yield f"{p}.attn.proj.weight", (VIT_DIM, VIT_DIM), "linear"
        yield f"{p}.attn.proj.bias", (VIT_DIM,), "bias"
        yield f"{p}.ls1.gamma", (VIT_DIM,), "ls"
        yield f"{p}.norm2.weight", (VIT_DIM,), "ln_w"
        yield f"{p}.norm2.bias", (VIT_DIM,), "ln_b"
        yield f"{p}.mlp.fc1.weight", (VIT_MLP, VIT_DIM), "linear"
        yield f"{p}.mlp.fc1.bias", (VIT_MLP,), "bias"
        yield f"{p}.mlp.fc2.weight", (VIT_DIM, VIT_MLP), "linear"
        yield f"{p}.mlp.fc2.bias", (VIT_DIM,), "bias"
        yield f"{p}.ls2.gamma", (VIT_DIM,), "ls"
    yield "norm.weight", (VIT_DIM,), "ln_w"
    yield "norm.bias", (VIT_DIM,), "ln_b"
