"""Two 4x128^3 forwards of the 94M `anatomix-dev` network (seeded init) for ncu captures:
    ncu --set full --clock-control none -k regex:<kernel> --launch-skip <n> --launch-count <m> -o gpurun_out/x python tools/ncu_forward_94m.py
"""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anatomix_b200 import Unet
from anatomix_b200.engine import Engine

cfg = dict(dimension=3, input_nc=1, output_nc=32, num_downs=5, ngf=32, norm="instance", pooling="Avg",
           interp="trilinear", norm_eps=1e-2)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    sd = Unet(**cfg).state_dict()
eng = Engine(cfg, "cuda:0")
eng.load_state(sd)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x = torch.rand(batch, 1, 128, 128, 128, device="cuda")
out = torch.empty((batch, 32, 128, 128, 128), device="cuda")
for _ in range(2):
    eng.forward(x, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
