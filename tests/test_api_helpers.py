"""The pure-Python helpers of the path (to_pixel_coordinates, to_normalized_coordinates, match_keypoints,
conf_from_fb_consistency; romatch/models/matcher.py:672-773) against vectors generated from the unmodified reference
(tests/golden/make_golden_helpers.py).  They never touch the engine, so they run on CPU on a bare instance."""
import os

import numpy as np
import pytest
import torch

from roma_b200.matcher import RegressionMatcher

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "helpers.npz"))


def _t(name):
    return torch.from_numpy(G[name])


def _model():
    return RegressionMatcher.__new__(RegressionMatcher)


def test_pixel_and_normalized_coordinates():
    m = _model()
    coords = _t("coords")
    pa, pb = m.to_pixel_coordinates(coords, 480, 640, 600, 800)
    assert torch.equal(pa, _t("pix_A")) and torch.equal(pb, _t("pix_B"))
    assert torch.equal(m.to_pixel_coordinates(coords[:, :2], 480, 640), _t("pix_single"))
    with pytest.raises(AttributeError):     # like the reference: `.shape` is read before the tuple branch (matcher.py:702)
        m.to_pixel_coordinates((coords[:, :2], coords[:, 2:]), 480, 640, 600, 800)
    na_t, nb_t = m.to_normalized_coordinates((pa, pb), 480, 640, 600, 800)                      # tuple form works here
    na, nb = m.to_normalized_coordinates(torch.cat((pa, pb), dim=-1), 480, 640, 600, 800)
    assert torch.equal(na, _t("norm_A")) and torch.equal(nb, _t("norm_B"))
    assert torch.equal(na_t, na) and torch.equal(nb_t, nb)
    assert (na - coords[:, :2]).abs().max() < 1e-6 and (nb - coords[:, 2:]).abs().max() < 1e-6   # round trip


def test_match_keypoints():
    m = _model()
    W = G["warp"].shape[1] // 2
    warp, cert = _t("warp")[:, :W], _t("certainty")[:, :W]
    x_A, x_B = _t("x_A"), _t("x_B")
    ia, ib = m.match_keypoints(x_A, x_B, warp, cert, return_tuple=True, return_inds=True, max_dist=0.005, cert_th=0.2)
    assert torch.equal(ia, _t("kp_inds_A")) and torch.equal(ib, _t("kp_inds_B")) and len(ia) > 10
    cat = m.match_keypoints(x_A, x_B, warp, cert, return_tuple=False, return_inds=False, max_dist=0.005, cert_th=0.2)
    assert torch.equal(cat, _t("kp_cat"))
    ka, kb = m.match_keypoints(x_A, x_B, warp, cert, max_dist=0.005, cert_th=0.2)
    assert torch.equal(torch.cat((ka, kb), dim=-1), cat)
    inds = m.match_keypoints(x_A, x_B, warp, cert, return_tuple=False, return_inds=True, max_dist=0.005, cert_th=0.2)
    assert torch.equal(inds, torch.cat((ia, ib), dim=-1))
    # nothing passes an impossible certainty threshold
    assert len(m.match_keypoints(x_A, x_B, warp, cert, return_inds=True, cert_th=2.0)[0]) == 0


def test_conf_from_fb_consistency():
    m = _model()
    W = G["warp"].shape[1] // 2
    a_to_b, b_to_a = _t("warp")[:, :W, 2:], _t("warp")[:, W:, :2]
    fb = m.conf_from_fb_consistency(a_to_b, b_to_a, th=2)
    assert fb.shape == (24, 32) and torch.equal(fb, _t("fb"))
    xs = torch.linspace(-1 + 1 / W, 1 - 1 / W, W)
    ys = torch.linspace(-1 + 1 / 24, 1 - 1 / 24, 24)
    grid = torch.stack(torch.meshgrid(xs, ys, indexing="xy"), dim=-1)
    fb2 = m.conf_from_fb_consistency(torch.stack((a_to_b, grid)), torch.stack((b_to_a, grid)), th=1)
    assert fb2.shape == (2, 24, 32) and torch.equal(fb2, _t("fb_batched"))
    assert fb2[1].min() == 1.0                          # the identity flow is consistent with itself everywhere


def test_visualize_warp(tmp_path):
    """visualize_warp (matcher.py:936-989) against the reference's output: PIL inputs (symmetric / one direction), tensor inputs,
    and the PNG written through `save_path` (tensor_to_pil, utils.py)."""
    from PIL import Image
    from roma_b200 import synthetic
    V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "helpers_visualize.npz"))
    m = _model()
    warp, cert = _t("warp"), _t("certainty")
    W = warp.shape[1] // 2
    im_a, im_b = synthetic.make_pil_pair(11, size_a=(50, 40), size_b=(45, 60))
    assert torch.equal(m.visualize_warp(warp, cert, im_a, im_b, device="cpu"), torch.from_numpy(V["vis_sym"]))
    assert torch.equal(m.visualize_warp(warp[:, :W], cert[:, :W], im_a, im_b, device="cpu", symmetric=False), torch.from_numpy(V["vis_one"]))
    path = str(tmp_path / "v.png")
    vis = m.visualize_warp(warp, cert, torch.from_numpy(V["x_A"]), torch.from_numpy(V["x_B"]), device="cpu", save_path=path)
    assert torch.equal(vis, torch.from_numpy(V["vis_tensor"]))
    assert np.array_equal(np.asarray(Image.open(path)), V["saved_png"])
