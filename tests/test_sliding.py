"""Sliding-window extraction: window grid / blend arithmetic against an independent
per-voxel CPU implementation (MONAI itself is not available offline: parity unpinned)."""
import math

import numpy as np
import pytest
import torch

from anatomix_b200.sliding import importance_map, scan_intervals, sliding_window_features, window_starts


def test_grid_matches_the_registration_setting():
    # 256^3 scan, 128^3 windows, overlap 0.8 -> stride 25 -> 7 windows per axis, last one pulled back to 128
    iv = scan_intervals((256, 256, 256), (128, 128, 128), 0.8)
    assert iv == [25, 25, 25]
    st = window_starts((256, 256, 256), (128, 128, 128), iv)
    assert len(st) == 7 ** 3 and st[0] == (0, 0, 0) and st[-1] == (128, 128, 128)
    assert sorted({s[0] for s in st}) == [0, 25, 50, 75, 100, 125, 128]
    assert scan_intervals((128, 200, 128), (128, 128, 128), 0.7) == [128, 38, 128]


def test_importance_map():
    g = importance_map((8, 8, 8), "gaussian", 0.25)
    assert g.shape == (8, 8, 8) and torch.isclose(g.max(), g[3, 3, 3]) and g.min() >= 1e-3
    x = torch.arange(-3.5, 4.5)
    one = torch.exp(-x * x / (2 * 2.0 * 2.0))
    assert torch.allclose(g[:, 3, 3] / g[3, 3, 3], one / one[3], atol=1e-6)
    assert torch.equal(importance_map((4, 4, 4), "constant"), torch.ones(4, 4, 4))


def brute_force(x, roi, overlap, mode, sigma_scale, fn):
    """Voxel-by-voxel accumulation, written independently of the module under test."""
    size = x.shape[2:]
    axes = []
    for img, r in zip(size, roi):
        step = r if r == img else max(int(r * (1 - overlap)), 1)
        n = int(math.ceil((img - r) / step)) + 1 if img > r else 1
        axes.append([min(i * step, img - r) for i in range(n)])
    if mode == "gaussian":
        ws = []
        for r in roi:
            c = (r - 1) / 2.0
            ws.append(np.exp(-((np.arange(r) - c) ** 2) / (2 * (sigma_scale * r) ** 2)))
        w = ws[0][:, None, None] * ws[1][None, :, None] * ws[2][None, None, :]
        w = np.maximum(w, max(w[w != 0].min(), 1e-3)).astype(np.float32)
    else:
        w = np.ones(roi, np.float32)
    acc, cnt = None, np.zeros(size, np.float32)
    for z in axes[0]:
        for y in axes[1]:
            for xx in axes[2]:
                p = fn(x[:, :, z:z + roi[0], y:y + roi[1], xx:xx + roi[2]]).numpy()[0]
                if acc is None:
                    acc = np.zeros((p.shape[0],) + tuple(size), np.float32)
                acc[:, z:z + roi[0], y:y + roi[1], xx:xx + roi[2]] += p * w
                cnt[z:z + roi[0], y:y + roi[1], xx:xx + roi[2]] += w
    return acc / cnt


@pytest.mark.parametrize("mode,overlap", [("gaussian", 0.8), ("constant", 0.7)])
def test_blend_matches_brute_force(mode, overlap):
    fn = lambda t: torch.cat([t * 2.0 + 1.0, t.flip(-1) - t.mean(dim=(2, 3, 4), keepdim=True)], dim=1)
    x = torch.rand(1, 1, 20, 17, 24, generator=torch.Generator().manual_seed(2))
    got = sliding_window_features(x, (8, 8, 16), 3, fn, overlap=overlap, mode=mode, sigma_scale=0.25)
    want = brute_force(x, (8, 8, 16), overlap, mode, 0.25, fn)
    np.testing.assert_allclose(got.numpy()[0], want, rtol=1e-5, atol=1e-5)


def test_small_volume_is_padded_and_cropped():
    fn = lambda t: t + 1.0
    x = torch.rand(2, 1, 6, 8, 5)
    y = sliding_window_features(x, (8, 8, 8), 4, fn, overlap=0.5, mode="gaussian", sigma_scale=0.25)
    assert y.shape == (2, 1, 6, 8, 5) and torch.allclose(y, x + 1.0, atol=1e-6)


def test_literal_fixture_of_the_monai_algorithm():
    """Pin against numbers derived BY HAND from MONAI's published `sliding_window_inference` algorithm (MONAI itself is
    not importable offline; reference call sites convex_adam_utils.py:202-219, train_segmentation.py:196-199), not
    against this module's own helpers.

    Image (4, 5, 6), window (2, 3, 4), overlap 0.5: scan interval per axis = int(roi * (1 - overlap)) = (1, 1, 2);
    windows per axis = ceil((img - roi) / interval) + 1 = (3, 3, 2), the last one pulled back to end at the border.
    Gaussian importance map, sigma = 0.25 * roi per axis, samples at -(n-1)/2 ... (n-1)/2:
      n = 2, sigma 0.5 : exp(-0.25 / 0.5)    = 0.60653066 (both taps)
      n = 3, sigma 0.75: exp(-1 / 1.125)     = 0.41111229, 1, 0.41111229
      n = 4, sigma 1.0 : exp(-2.25 / 2)      = 0.32465247, exp(-0.25 / 2) = 0.88249690 (mirrored)
    separable product, floored at max(min, 1e-3) = its own minimum 0.60653066 * 0.41111229 * 0.32465247 = 0.08095281."""
    from anatomix_b200.sliding import importance_map, scan_intervals, window_starts
    assert scan_intervals((4, 5, 6), (2, 3, 4), 0.5) == [1, 1, 2]
    starts = window_starts((4, 5, 6), (2, 3, 4), [1, 1, 2])
    assert starts == [(z, y, x) for z in (0, 1, 2) for y in (0, 1, 2) for x in (0, 2)]
    w = importance_map((2, 3, 4), "gaussian", 0.25)
    gz, gy, gx = [0.60653066, 0.60653066], [0.41111229, 1.0, 0.41111229], [0.32465247, 0.88249690, 0.88249690, 0.32465247]
    want = torch.tensor([[[a * b * c for c in gx] for b in gy] for a in gz])
    assert torch.allclose(w, want, atol=1e-7, rtol=1e-6) and abs(w.min().item() - 0.08095281) < 1e-7
    # the registration setting (256^3 scan, 128^3 windows, overlap 0.8): interval int(128 * 0.2) = 25, seven
    # windows per axis, the last pulled back from 150 to 128 -> 343 windows
    assert scan_intervals((256,) * 3, (128,) * 3, 0.8) == [25, 25, 25]
    s = window_starts((256,) * 3, (128,) * 3, [25, 25, 25])
    assert len(s) == 343 and sorted({a for a, _, _ in s}) == [0, 25, 50, 75, 100, 125, 128]
    # a window as large as the image along an axis: one window there, interval = the whole axis
    assert scan_intervals((8, 6), (8, 3), 0.5) == [8, 1] and window_starts((8, 6), (8, 3), [8, 1]) == [(0, 0), (0, 1), (0, 2), (0, 3)]
    # blended output of a constant predictor is that constant everywhere (weights normalise out), whatever the overlap
    from anatomix_b200.sliding import sliding_window_features
    x = torch.zeros(1, 1, 4, 5, 6)
    y = sliding_window_features(x, (2, 3, 4), 3, lambda p: torch.full((p.shape[0], 2) + tuple(p.shape[2:]), 7.0), overlap=0.5,
                                mode="gaussian", sigma_scale=0.25)
    assert y.shape == (1, 2, 4, 5, 6) and torch.allclose(y, torch.full_like(y, 7.0), atol=1e-6)
