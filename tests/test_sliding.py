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
