"""Multi-rank host logic on CPU: world_size-2 gloo process groups (no GPU).
Batch sharding + feature all-gather with the oracle as the compute function,
ragged shards, depth-slab bounds and the halo-plane exchange."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from anatomix_b200.dist import ShardedExtractor, exchange_halo_planes, shard_range, shard_sizes, slab_bounds


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_slab_bounds_alignment():
    b = slab_bounds(512, 8, 4)
    assert b[0] == (0, 64) and b[-1] == (448, 512)
    b = slab_bounds(160, 2, 4)           # 10 blocks of 16 -> 5 + 5
    assert b == [(0, 80), (80, 160)]
    b = slab_bounds(112, 3, 4)           # 7 blocks -> 3, 2, 2
    assert [hi - lo for lo, hi in b] == [48, 32, 32]
    with pytest.raises(ValueError):
        slab_bounds(100, 2, 4)           # not a multiple of 16
    with pytest.raises(ValueError):
        slab_bounds(48, 2, 4)            # slabs would be thinner than 32 planes


def _worker(rank, world, port, ragged):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        from oracle import unet_oracle as O
        cfg = dict(dimension=3, input_nc=1, output_nc=8, num_downs=1, ngf=8)
        state = O.random_state(cfg, seed=1)
        compute = lambda t: O.unet_forward(cfg, state, t)
        n = 3 if ragged else 4
        batch = torch.rand(n, 1, 8, 8, 8, generator=torch.Generator().manual_seed(5))
        ex = ShardedExtractor(compute, 8)
        mine = ex.extract(batch)
        lo, hi = shard_range(n, world, rank)
        want = compute(batch)
        assert torch.allclose(mine, want[lo:hi], atol=1e-6)
        full = ex.extract(batch, gather=True)
        assert full.shape == want.shape and torch.allclose(full, want, atol=1e-6)
        full2 = ex.extract(batch[lo:hi], gather=True, presharded=True, n_total=n)
        assert torch.allclose(full2, want, atol=1e-6)

        # halo exchange on a depth-split tensor with a one-plane shell
        vol = torch.arange(2 * 3 * 8 * 5, dtype=torch.float32).view(2, 3, 8, 5)     # [n, g, D, plane]
        lo_z, hi_z = (0, 4) if rank == 0 else (4, 8)
        slab = torch.full((2, 3, 6, 5), -1.0)
        slab[:, :, 1:5] = vol[:, :, lo_z:hi_z]
        lo_in, hi_in = torch.empty(2, 3, 5), torch.empty(2, 3, 5)
        exchange_halo_planes(slab[:, :, 1].contiguous(), slab[:, :, 4].contiguous(), lo_in, hi_in, rank, world)
        if rank == 0:
            assert torch.equal(hi_in, vol[:, :, 4])
        else:
            assert torch.equal(lo_in, vol[:, :, 3])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_two_rank_gloo_sharding_and_halo(ragged):
    port = 29500 + (os.getpid() % 2000) + (1 if ragged else 0)
    mp.spawn(_worker, args=(2, port, ragged), nprocs=2, join=True)
