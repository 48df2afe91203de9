import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import unet_oracle as O
from anatomix_b200.engine import Engine
cfg = dict(dimension=3, input_nc=1, output_nc=16, num_downs=2, ngf=16)
state = O.random_state(cfg, seed=5)
x = torch.rand(1, 1, 8, 16, 8).cuda()
eng = Engine(cfg, "cuda:0"); eng.load_state(state)
out = torch.empty(1, 16, 8, 16, 8, device="cuda")
for i, st in enumerate(eng.step_table()):
    eng.run_steps(x, out, i, i + 1)
    try:
        torch.cuda.synchronize()
        print("step", i, st[-1], "ok")
    except Exception as e:
        print("step", i, st[-1], "FAILED", str(e)[:100]); break
