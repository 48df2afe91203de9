"""Two 8x128^3 forwards of the 6M network (the second one is the one to capture under ncu):
    ncu --set full --clock-control none --import-source on -k regex:'conv3_umma|conv3_rows|stem_umma' --launch-skip 21 --launch-count 21 \
        -o gpurun_out/fwd python tools/ncu_forward.py
"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anatomix_b200.engine import Engine

cfg = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
z = np.load(os.path.join(ROOT, "tests/golden/anatomix_6m_state.npz"))
eng = Engine(cfg, "cuda:0")
eng.load_state({k: torch.from_numpy(z[k]) for k in z.files})
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x = torch.rand(batch, 1, 128, 128, 128, device="cuda")
out = torch.empty((batch, 16, 128, 128, 128), device="cuda")
for _ in range(2):
    eng.forward(x, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
