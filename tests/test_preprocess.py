"""Image preprocessing: the Pillow restatement in oracle/ against Pillow itself, the library's host
coefficient routine against the restatement, and the CUDA path against both -- all bit-exact (integer work; the float
stage keeps the reference's fp32 operation order, `romatch/utils/utils.py:164-183,250-260`)."""
import numpy as np
import pytest
import torch
from PIL import Image

from oracle import pil_resample
from roma_b200 import preprocess

SIZES = [  # (H, W) -> (h, w)
    ((150, 200), (112, 112)),      # the PIL-route golden's image A
    ((220, 180), (168, 168)),
    ((97, 131), (224, 168)),       # up-sampling both ways
    ((480, 640), (560, 560)),      # down in x, up in y
    ((700, 500), (112, 168)),
    ((64, 64), (64, 128)),         # height unchanged: vertical pass skipped
    ((64, 80), (32, 80)),          # width unchanged: horizontal pass skipped
    ((56, 70), (56, 70)),          # identity
    ((33, 47), (47, 33)),
]


def _image(shape, seed):
    rng = np.random.default_rng(seed)
    if seed % 2:                                   # smooth blobs + hard edges (exercises clip8 on overshoot)
        img = np.zeros(shape + (3,), np.uint8)
        img[shape[0] // 4: shape[0] // 2, shape[1] // 3:] = 255
        img[::7, :, 1] = 255
        return img
    return rng.integers(0, 256, shape + (3,), dtype=np.uint8)


@pytest.mark.parametrize("idx", range(len(SIZES)))
def test_oracle_resize_is_pillow(idx):
    (H, W), (h, w) = SIZES[idx]
    img = _image((H, W), idx)
    ref = np.asarray(Image.fromarray(img, "RGB").resize((w, h), Image.BICUBIC))
    assert np.array_equal(pil_resample.resize_bicubic_u8(img, (h, w)), ref)
    # the float stage: the oracle against the host statement of the reference transform
    want = preprocess.pil_to_normalized(Image.fromarray(img, "RGB"), (h, w)).numpy()
    assert np.array_equal(pil_resample.preprocess(img, (h, w)), want)


def test_oracle_large_downscale_is_pillow():
    img = _image((1200, 1600), 4)
    ref = np.asarray(Image.fromarray(img, "RGB").resize((112, 168), Image.BICUBIC))
    assert np.array_equal(pil_resample.resize_bicubic_u8(img, (168, 112)), ref)


@pytest.mark.parametrize("n_in,n_out", [(200, 112), (131, 168), (4000, 560), (560, 864), (33, 47), (64, 64), (3, 100), (1000, 3)])
def test_library_coefficients_match_oracle(n_in, n_out):
    ks, bounds, kk = preprocess.resample_coeffs(n_in, n_out)      # host-only entry point of the C ABI
    ks2, bounds2, kk2 = pil_resample.coeffs(n_in, n_out)
    assert ks == ks2 and np.array_equal(bounds, bounds2) and np.array_equal(kk, kk2)
    assert (kk.sum(axis=1) - (1 << 22)).__abs__().max() <= ks      # rows sum to 1.0 in 22-bit fixed point up to rounding


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(SIZES)))
def test_device_preprocess_is_pillow(idx):
    (H, W), (h, w) = SIZES[idx]
    img = _image((H, W), idx)
    pil = Image.fromarray(img, "RGB")
    pre = preprocess.DevicePreprocessor("cuda")
    raw = pre.upload(pil)
    u8 = torch.empty(h, w, 3, dtype=torch.uint8, device="cuda")
    out = pre.resize_normalize(raw, (h, w), out_u8=u8)
    ref_u8 = np.asarray(pil.resize((w, h), Image.BICUBIC))
    assert np.array_equal(u8.cpu().numpy(), ref_u8)
    assert torch.equal(out.cpu(), preprocess.pil_to_normalized(pil, (h, w)))
    assert np.array_equal(out.cpu().numpy(), pil_resample.preprocess(img, (h, w)))


@pytest.mark.gpu
def test_device_preprocess_photo_size():
    """A 12-megapixel frame down to the two network resolutions from one upload."""
    img = _image((3000, 4000), 6)
    pil = Image.fromarray(img, "RGB")
    pre = preprocess.DevicePreprocessor("cuda")
    raw = pre.upload(pil)
    for size in ((560, 560), (864, 864)):
        out = pre.resize_normalize(raw, size)
        assert torch.equal(out.cpu(), preprocess.pil_to_normalized(pil, size))


@pytest.mark.gpu
def test_device_preprocess_rejects_bad_arguments():
    pre = preprocess.DevicePreprocessor("cuda")
    with pytest.raises(NotImplementedError):
        pre.upload(Image.new("L", (8, 8)))
    raw = torch.zeros(8, 8, 3, dtype=torch.uint8, device="cuda")
    from roma_b200 import cabi
    with pytest.raises(RuntimeError):              # resized width without tables
        cabi.call("romab200_preprocess_rgb8", "rb_preprocess_args", ld_in=24, in_h=8, in_w=8, out_h=8, out_w=16,
                  out=torch.empty(3, 8, 16, device="cuda"), mean=[0, 0, 0], std=[1, 1, 1], **{"in": raw})
