"""Sliding-window feature extraction on the engine.

The reference's production caller runs the U-Net through MONAI's
``sliding_window_inference`` (registration: 128^3 windows, ``sw_batch_size=2``,
``overlap=0.8``, gaussian weights, ``sigma_scale=0.25`` --
``anatomix/registration/convex_adam_utils.py:202-219``; segmentation validation:
``overlap=0.7``, constant weights -- ``anatomix/segmentation/train_segmentation.py:196-199``).
This module restates that algorithm (window grid, importance map, weighted
accumulate / normalise) for GPU-resident volumes so whole scans go through the
engine in large window batches.

MONAI (an unpinned dependency of the reference, ``requirements.txt:12``) is not
available offline, so this restatement is pinned against a literal fixture derived
BY HAND from MONAI's published algorithm (window origins, scan intervals and the
gaussian importance map of a tiny case plus the registration setting's 343-window
grid: ``tests/test_sliding.py::test_literal_fixture_of_the_monai_algorithm``) and
against an independent CPU implementation, not against MONAI's code itself.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn.functional as F


_LIB = None


def _lib_or_none():
    """The engine library for the native blend kernel (CUDA tensors only; torch ops otherwise)."""
    global _LIB
    if _LIB is None:
        try:
            from . import _lib
            _LIB = _lib.load()
        except Exception:
            _LIB = False
    return _LIB or None


def scan_intervals(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> List[int]:
    """Per-axis stride between windows: the whole axis when the window covers it,
    else ``int(roi * (1 - overlap))`` (at least 1)."""
    out = []
    for img, roi in zip(image_size, roi_size):
        if roi == img:
            out.append(int(roi))
        else:
            step = int(roi * (1.0 - overlap))
            out.append(step if step > 0 else 1)
    return out


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], intervals: Sequence[int]) -> List[Tuple[int, ...]]:
    """Window origins, first axis slowest; the last window of an axis is pulled back
    so it ends at the border."""
    per_axis = []
    for img, roi, step in zip(image_size, roi_size, intervals):
        num = int(math.ceil(float(img - roi) / step)) + 1 if img > roi else 1
        starts = []
        for i in range(num):
            s = i * step
            s -= max(s + roi - img, 0)
            starts.append(s)
        per_axis.append(starts)
    return list(itertools.product(*per_axis))


def importance_map(roi_size: Sequence[int], mode: str = "gaussian", sigma_scale: float = 0.25,
                   device="cpu") -> torch.Tensor:
    """Blend weights of one window: ones (``constant``) or a separable gaussian with
    ``sigma = sigma_scale * roi`` centred on the window, floored at its smallest
    non-zero value (>= 1e-3) so no voxel gets zero weight."""
    if mode == "constant":
        return torch.ones(tuple(roi_size), dtype=torch.float32, device=device)
    if mode != "gaussian":
        raise ValueError(f"unsupported blend mode {mode!r}")
    w = None
    for n in roi_size:
        sigma = sigma_scale * n
        x = torch.arange(-(n - 1) / 2.0, (n - 1) / 2.0 + 1, dtype=torch.float32, device=device)
        g = torch.exp(-(x * x) / (2.0 * sigma * sigma))
        w = g if w is None else w.unsqueeze(-1) * g
    floor = max(w[w != 0].min().item(), 1e-3)
    return w.clamp(min=floor)


@torch.no_grad()
def sliding_window_features(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                            predictor: Callable[[torch.Tensor], torch.Tensor], overlap: float = 0.25,
                            mode: str = "constant", sigma_scale: float = 0.125) -> torch.Tensor:
    """``[B, C, D, H, W] -> [B, C_out, D, H, W]`` by blending ``predictor`` outputs of
    overlapping ``roi_size`` windows (argument order of MONAI's ``sliding_window_inference``).
    Volumes smaller than the window are zero-padded symmetrically and cropped back."""
    roi = tuple(int(r) for r in roi_size)
    orig = tuple(inputs.shape[2:])
    pads = []
    for img, r in zip(reversed(orig), reversed(roi)):          # F.pad order: last axis first
        diff = max(r - img, 0)
        pads += [diff // 2, diff - diff // 2]
    x = F.pad(inputs, pads, mode="constant", value=0.0) if any(pads) else inputs
    size = tuple(x.shape[2:])
    starts = window_starts(size, roi, scan_intervals(size, roi, overlap))
    weight = importance_map(roi, mode, sigma_scale, device=x.device)
    batch = x.shape[0]
    windows = [(b,) + s for b in range(batch) for s in starts]
    out = None
    norm = torch.zeros((batch, 1) + size, dtype=torch.float32, device=x.device)
    for i in range(0, len(windows), sw_batch_size):
        group = windows[i:i + sw_batch_size]
        patch = torch.stack([x[b, :, z:z + roi[0], y:y + roi[1], w:w + roi[2]] for b, z, y, w in group])
        pred = predictor(patch.contiguous())
        if out is None:
            out = torch.zeros((batch, pred.shape[1]) + size, dtype=torch.float32, device=x.device)
        native = pred.is_cuda and pred.dtype == torch.float32 and _lib_or_none() is not None
        if native:
            pred = pred.contiguous()
            stream = torch.cuda.current_stream(pred.device).cuda_stream
        for k, (b, z, y, w) in enumerate(group):
            if native:      # one streaming pass per window: weight read once, no temporaries
                with torch.cuda.device(pred.device):
                    st = _LIB.anx_blend_window_f32(pred[k].data_ptr(), weight.data_ptr(), out[b].data_ptr(),
                                                   norm[b].data_ptr(), pred.shape[1], roi[0], roi[1], roi[2],
                                                   size[0], size[1], size[2], z, y, w, stream)
                if st != 0:
                    raise RuntimeError(f"anx_blend_window_f32 failed with status {st}")
                continue
            out[b, :, z:z + roi[0], y:y + roi[1], w:w + roi[2]] += pred[k] * weight
            norm[b, :, z:z + roi[0], y:y + roi[1], w:w + roi[2]] += weight
    out /= norm
    if any(pads):
        sl = [slice(None), slice(None)]
        for axis, (img, r) in enumerate(zip(orig, roi)):
            lo = max(r - img, 0) // 2
            sl.append(slice(lo, lo + img))
        out = out[tuple(sl)]
    return out
