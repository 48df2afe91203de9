"""Measurements of the rows next to the hot path, one JSON object on stdout:
  * sliding-window feature extraction of one 256^3 scan (registration setting: 128^3 windows, overlap 0.8,
    gaussian blend) fed to the engine in window batches of 8;
  * forward with feature taps (the four encoder skip tensors) vs the plain forward;
  * forward with a fused 5-class 1x1x1 head vs plain forward + cuDNN pointwise conv;
  * scale + average pooling kernel: achieved GB/s against the measured HBM copy peak.
Usage on a GPU box:  python tools/bench_rows.py"""
import contextlib, io, json, os, sys
import numpy as np
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anatomix_b200 import Unet
from anatomix_b200.engine import Engine
from anatomix_b200.heads import UnetOutBlock, avg_pool3d_scaled, fuse_output_head
from anatomix_b200.sliding import sliding_window_features

CFG = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
z = np.load(os.path.join(ROOT, "tests/golden/anatomix_6m_state.npz"))
state = {k: torch.from_numpy(z[k]) for k in z.files}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


res = {}
with contextlib.redirect_stdout(io.StringIO()):
    m = Unet(**CFG)
m.load_state_dict(state)
m = m.cuda().eval()
x = torch.rand(8, 1, 128, 128, 128, device="cuda")
with torch.no_grad():
    plain = timed(lambda: m(x))
    res["forward_8x128_ms"] = plain
    # taps: the four skip tensors (encoder_idx) next to the output
    res["forward_with_4_skip_taps_ms"] = timed(lambda: m(x, layers=list(m.encoder_idx)))
    res["encode_only_to_skip3_ms"] = timed(lambda: m(x, layers=list(m.encoder_idx), encode_only=True))
    # fused segmentation head
    torch.manual_seed(0)
    head = UnetOutBlock(3, 16, 5, False).cuda()
    seq = fuse_output_head(m, head).cuda().eval()
    res["fused_head5_ms"] = timed(lambda: seq(x))
    res["unfused_head5_ms"] = timed(lambda: head(m(x)))
    res["fused_head5_volumes_per_s"] = 8e3 / res["fused_head5_ms"]
    # scale + average pooling of the feature volume
    feats = m(x)
    for k in (2, 4):
        ms = timed(lambda: avg_pool3d_scaled(feats, k, 0.1))
        ms_t = timed(lambda: F.avg_pool3d(feats * 0.1, k, stride=k))
        moved = feats.numel() * 4 * (1 + 1 / k ** 3)
        res[f"avgpool_k{k}"] = {"ms": ms, "torch_ms": ms_t, "gbs": moved / ms / 1e6, "hbm_peak_gbs": peaks["hbm_gbs"],
                                "frac": moved / ms / 1e6 / peaks["hbm_gbs"]}
    del feats
    # sliding-window scan (convex_adam_utils.py:202-219 setting), window batches of 8
    scan = torch.rand(1, 1, 256, 256, 256, device="cuda")
    fn = lambda: sliding_window_features(scan, (128, 128, 128), 8, m, overlap=0.8, mode="gaussian", sigma_scale=0.25)
    ms = timed(fn, steps=2, warmup=1)
    res["sliding_256_overlap0.8"] = {"ms": ms, "windows": 343, "windows_per_s": 343e3 / ms,
                                     "engine_only_ms_estimate": 343 / 8 * plain}
print(json.dumps(res))
