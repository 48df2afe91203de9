"""Timing experiments on the conv kernel (results are WRONG under ablation; this
only locates the bottleneck).  The switches exist only in experiment builds of the library:
    ANX_LIB_VARIANT=exp python -m anatomix_b200.build          (here, before shipping the tree to the GPU box)
    ANX_LIB_VARIANT=exp python tools/ablate.py                 (on the GPU box)"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, os, torch, numpy as np
sys.path.insert(0, %r)
from anatomix_b200.engine import Engine
cfg = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
z = np.load(os.path.join(%r, "tests/golden/anatomix_6m_state.npz"))
eng = Engine(cfg, "cuda:0"); eng.load_state({k: torch.from_numpy(z[k]) for k in z.files})
x = torch.rand(4, 1, 128, 128, 128, device="cuda")
for _ in range(2): eng.forward(x)
acc = {}
for r in range(3):
    for n, t in eng.profile(x): acc[n] = acc.get(n, 0) + t / 3
print(" ".join(f"{n.split('_')[0]}:{acc[n]*1000:.0f}" for n in acc if n.startswith("conv")))
''' % (ROOT, ROOT)
if not os.environ.get("ANX_LIB_VARIANT"):
    sys.exit("set ANX_LIB_VARIANT=<tag> (an -DANX_EXPERIMENTS build): the product library ignores ANX_ABLATE")
for ab in sys.argv[1:] or ["0", "1", "2", "4", "8", "3", "5", "6", "15"]:
    env = dict(os.environ, ANX_ABLATE=ab)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(f"ablate={ab:>2} us:", r.stdout.strip() or r.stderr[-400:])
