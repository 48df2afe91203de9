"""Informational: the reference's own op sequence (stock torch modules -> cuDNN / ATen CUDA kernels) on the
same B200, same network, same batch (last row).  This is NOT the reference arm of
bench.py (that one is the reference's CPU path) -- it is the "recompiled library kernels" baseline the
engine is meant to beat: fp32, TF32, and bf16 autocast with channels_last_3d.
Usage on a GPU box:  python tools/bench_torch_gpu.py [batch]      -> one JSON object on stdout"""
import contextlib, io, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ANATOMIX_B200_DISABLE"] = "1"          # the module walks its stock nn.Sequential
from anatomix_b200 import Unet

CFG_6M = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
with contextlib.redirect_stdout(io.StringIO()):
    m = Unet(**CFG_6M)
z = np.load(os.path.join(ROOT, "tests/golden/anatomix_6m_state.npz"))
m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
m = m.cuda().eval()
x = torch.rand(batch, 1, 128, 128, 128, device="cuda")


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


res = {"batch": batch, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
       "cudnn": torch.backends.cudnn.version()}
with torch.no_grad():
    ref = m(x).float()
    for name, tf32, autocast, cl in (("fp32", False, False, False), ("tf32", True, False, False),
                                     ("bf16_autocast", True, True, False),
                                     ("bf16_autocast_channels_last_3d", True, True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        mm = m.to(memory_format=torch.channels_last_3d) if cl else m.to(memory_format=torch.contiguous_format)
        xx = x.contiguous(memory_format=torch.channels_last_3d) if cl else x
        def run():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                return mm(xx)
        try:
            ms = timed(run)
            y = run().float()
            rel = ((y - ref).norm() / ref.norm()).item()
            res[name] = {"ms_per_step": ms, "volumes_per_s": batch / ms * 1e3, "rel_l2_vs_fp32": rel}
        except Exception as ex:          # e.g. out of memory for a layout: report, keep going
            res[name] = {"error": repr(ex)[:200]}
        torch.cuda.empty_cache()
print(json.dumps(res))
