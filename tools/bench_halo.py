"""Depth-halo partition of ONE oversized volume over the GPUs of a box (BASELINE configs[4]):
    torchrun --nproc-per-node 8 tools/bench_halo.py [--size 512]
Checks the partitioned result against the single-GPU engine on a 256-plane volume, then times
the 1 x size^3 volume.  Prints one JSON line on rank 0."""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anatomix_b200.engine import Engine
from anatomix_b200.halo import DepthSlabExtractor
from anatomix_b200.dist import slab_bounds

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
z = np.load(os.path.join(ROOT, "tests/golden/anatomix_6m_state.npz"))
state = {k: torch.from_numpy(z[k]) for k in z.files}
slab = DepthSlabExtractor(cfg, state, dev)

# correctness: 32 planes per rank, against the single-GPU engine
depth = 32 * world
vol = torch.rand(1, 1, depth, 64, 64, generator=torch.Generator().manual_seed(11))
mine = slab.extract(vol)
lo, hi = slab_bounds(depth, world, 4)[rank]
ref = Engine(cfg, dev); ref.load_state(state)
want = ref.forward(vol.to(dev))[:, :, lo:hi]
torch.cuda.synchronize()
err = (mine - want).abs().max()
dist.all_reduce(err, op=dist.ReduceOp.MAX)
del ref, want

big = torch.rand(1, 1, a.size, a.size, a.size, generator=torch.Generator().manual_seed(12))
slab.extract(big)                       # warm-up (workspace, plans)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# time the engine part only: the host->device upload of the slab is identical for every approach
n, _, D, H, W = big.shape
lo, hi = slab_bounds(D, world, 4)[rank]
from anatomix_b200.halo import slab_input_with_halo
x = slab_input_with_halo(big, lo, hi).to(dev)
out = torch.empty((1, 16, hi - lo, H, W), device=dev)
ws = slab.engine.workspace(1, hi - lo, H, W)
table = slab.engine.buffer_table(1, hi - lo, H, W)
def run():
    for i, (kind, buf, goff, groups, name) in enumerate(slab.steps):
        slab.engine.run_steps(x, out, i, i + 1)
        if buf >= 0:
            slab._exchange(ws, table, buf, goff, groups, 1, hi - lo, H, W)
run()
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(a.steps):
    run()
e1.record()
dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": f"anatomix 6M UNet, one 1x{a.size}^3 volume, depth-halo partition over {world} GPUs",
                      "ms_per_volume": t.item(), "volumes_per_s": 1e3 / t.item(),
                      "equiv_128_volumes_per_s": (a.size / 128) ** 3 * 1e3 / t.item(),
                      "max_abs_err_vs_single_gpu": err.item(), "exchanges_per_forward": sum(1 for s in slab.steps if s[1] >= 0)}))
dist.destroy_process_group()
