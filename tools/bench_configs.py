"""Informational timings of the other BASELINE.json configs (not the headline bench line):
    94M `anatomix-dev` (seeded init), batch 4 x 128^3 on one GPU        (configs[2])
    6M, one 1 x 512^3 volume on one GPU (single-GPU reference for the depth-halo partition of configs[4])
Usage on a GPU box:  python tools/bench_configs.py [94m] [512]"""
import contextlib, io, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anatomix_b200 import Unet
from anatomix_b200.engine import Engine

CFG_6M = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
CFG_94M = dict(dimension=3, input_nc=1, output_nc=32, num_downs=5, ngf=32, norm="instance", pooling="Avg",
               interp="trilinear", norm_eps=1e-2)


def timed(eng, x, steps=5, warmup=2):
    out = torch.empty((x.shape[0], eng.output_nc) + tuple(x.shape[2:]), device=x.device)
    for _ in range(warmup):
        eng.forward(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.forward(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


which = sys.argv[1:] or ["94m", "512"]
res = {}
if "94m" in which:
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_94M)
    eng = Engine(CFG_94M, "cuda:0"); eng.load_state(m.state_dict())
    x = torch.rand(4, 1, 128, 128, 128, device="cuda")
    ms = timed(eng, x)
    prof = {}
    for n, t in eng.profile(x):
        prof[n] = round(t, 4)
    res["94m_4x128"] = {"ms_per_step": ms, "volumes_per_s": 4 / ms * 1e3, "tflops": 1418.748 * 4 / ms, "launch_ms": prof,
                        "workspace_gib": eng.workspace_bytes(4, 128, 128, 128) / 2**30}
    del eng
if "512" in which:
    z = np.load(os.path.join(ROOT, "tests/golden/anatomix_6m_state.npz"))
    eng = Engine(CFG_6M, "cuda:0"); eng.load_state({k: torch.from_numpy(z[k]) for k in z.files})
    x = torch.rand(1, 1, 512, 512, 512, device="cuda")
    ms = timed(eng, x, steps=3, warmup=1)
    res["6m_1x512"] = {"ms_per_step": ms, "volumes512_per_s": 1e3 / ms, "equiv_128_volumes_per_s": 64e3 / ms,
                       "workspace_gib": eng.workspace_bytes(1, 512, 512, 512) / 2**30}
print(json.dumps(res))
