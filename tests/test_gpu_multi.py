"""Two-GPU paths (NCCL): batch-sharded extraction with the feature all-gather and the
depth-halo partition of one volume, each against the single-GPU engine.  Skipped on
boxes with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import CFG_6M, GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _state():
    z = np.load(os.path.join(GOLDEN, "anatomix_6m_state.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from anatomix_b200.dist import ShardedExtractor
        from anatomix_b200.engine import Engine
        from anatomix_b200.halo import DepthSlabExtractor
        state = _state()
        eng = Engine(CFG_6M, dev)
        eng.load_state(state)

        # (1) batch sharding + all-gather of the 16-channel features
        batch = torch.rand(4, 1, 32, 32, 32, generator=torch.Generator().manual_seed(3))
        ex = ShardedExtractor(lambda t: eng.forward(t.to(dev)), 16)
        full = ex.extract(batch.to(dev), gather=True)
        ref = eng.forward(batch.to(dev))
        torch.cuda.synchronize()
        assert torch.equal(full, ref), "gathered features differ from the single-GPU result"

        # (1b) the same gather fused into the last conv's epilogue (peer stores over NVLink)
        from anatomix_b200.dist import FusedGatherExtractor, shard_range
        fused = FusedGatherExtractor(eng)
        lo, hi = shard_range(4, world, rank)
        got = fused.extract(batch[lo:hi].to(dev))
        torch.cuda.synchronize()
        assert torch.equal(got, ref), "fused gather differs from the single-GPU result"
        got2 = fused.extract(batch[lo:hi].to(dev) * 0.5)      # buffers are reused call after call
        torch.cuda.synchronize()
        assert torch.equal(got2, eng.forward(batch.to(dev) * 0.5))

        # (1c) copy-engine push gather (anx_push_to_peers), fp32 payload: bit-identical as well, buffers reused
        from anatomix_b200.dist import FeatureGather
        push = FeatureGather(eng, mode="push", payload="f32")
        for scale in (1.0, 0.5, 0.25):                      # three submissions: both buffers of the ring are reused
            got = push.extract(batch[lo:hi].to(dev) * scale)
            torch.cuda.synchronize()
            assert torch.equal(got, eng.forward(batch.to(dev) * scale)), "push gather differs from the single-GPU result"
        # pipelined use: two submissions in flight before the first is read
        s0 = push.submit(batch[lo:hi].to(dev))
        s1 = push.submit(batch[lo:hi].to(dev) * 2.0)
        g0 = push.wait(s0).clone()
        g1 = push.wait(s1).clone()
        torch.cuda.synchronize()
        assert torch.equal(g0, ref) and torch.equal(g1, eng.forward(batch.to(dev) * 2.0))

        # (1d) 16-bit channels-last payload, pushed and fused: equal to the local 16-bit forward of the whole batch,
        # and, widened, within one 16-bit rounding of the fp32 features
        local16 = eng.forward_cl16(batch.to(dev))
        for mode in ("push", "fused"):
            g16 = FeatureGather(eng, mode=mode, payload="cl16").extract(batch[lo:hi].to(dev))
            torch.cuda.synchronize()
            assert g16.dtype == eng.storage_dtype and tuple(g16.shape) == (4, 32, 32, 32, 16)
            assert torch.equal(g16, local16), f"{mode} cl16 gather differs from the local 16-bit forward"
            wide = eng.widen(g16)
            torch.cuda.synchronize()
            assert torch.equal(wide, ref.to(eng.storage_dtype).float()), "widened payload is not the rounded fp32 output"

        # (2) one volume, depth split in two slabs of 32 planes with per-layer halo planes: the in-engine
        # exchange (peer stores + flags, anx_engine_forward_slab) and the step-wise NCCL protocol
        vol = torch.rand(1, 1, 64, 32, 48, generator=torch.Generator().manual_seed(4))
        want = eng.forward(vol.to(dev))
        for in_engine in (True, False):
            slab = DepthSlabExtractor(CFG_6M, state, dev, in_engine_exchange=in_engine)
            assert slab.in_engine_exchange == in_engine
            for rep in range(2):                             # twice: sequence numbers carry over between forwards
                got = slab.extract(vol, gather=True)
                torch.cuda.synchronize()
                err = (got - want).abs().max().item()
                rel = ((got - want).norm() / want.norm()).item()
                with open(os.path.join(out_dir, f"rank{rank}.txt"), "a") as f:
                    f.write(f"slab in_engine={in_engine} rep={rep}: max abs {err} rel-L2 {rel}\n")
                assert err == 0.0, f"depth-slab result (in_engine={in_engine}) differs from single-GPU: max abs {err}, rel-L2 {rel}"
        # a 128-wide volume: the row kernel's launches take part in the exchange as well
        vol = torch.rand(1, 1, 64, 32, 128, generator=torch.Generator().manual_seed(6))
        want = eng.forward(vol.to(dev))
        got = DepthSlabExtractor(CFG_6M, state, dev).extract(vol, gather=True)
        torch.cuda.synchronize()
        assert torch.equal(got, want), "depth-slab result differs on the row-kernel shape"

        # (3) the same partition for an InstanceNorm / AvgPool / trilinear network (`anatomix-dev` style):
        # whole-volume statistics via an all-reduce of the per-conv sums, neighbour planes in the upsample
        from oracle import unet_oracle as O
        cfg_in = dict(dimension=3, input_nc=1, output_nc=16, num_downs=2, ngf=16, norm="instance",
                      pooling="Avg", interp="trilinear", norm_eps=1e-2)
        st_in = O.random_state(cfg_in, seed=9)
        vol = torch.rand(1, 1, 32, 16, 24, generator=torch.Generator().manual_seed(5))
        single = Engine(cfg_in, dev)
        single.load_state(st_in)
        want = single.forward(vol.to(dev))
        got = DepthSlabExtractor(cfg_in, st_in, dev).extract(vol, gather=True)
        torch.cuda.synchronize()
        rel = ((got - want).norm() / want.norm()).item()
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "a") as f:
            f.write(f"instance-norm slab: rel-L2 {rel}\n")
        # not bit-exact: the statistics are summed in a different order, which can flip a 16-bit rounding
        assert rel < 5e-3, f"InstanceNorm depth-slab result differs from single-GPU: rel-L2 {rel}"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharding_and_depth_halo(tmp_path):
    port = 29700 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        print(open(tmp_path / f"rank{r}.txt").read())
