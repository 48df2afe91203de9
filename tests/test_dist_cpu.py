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


def _slab_worker(rank, world, port):
    """The slab protocol (anatomix_b200.halo.run_slab_program) on a toy two-layer network executed with torch
    ops: z-conv -> whole-volume normalisation -> z-conv -> normalisation, depth split over two ranks with a
    one-plane shell, against the unsplit computation."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from anatomix_b200.halo import run_slab_program
        D, P = 8, 6                                            # depth, voxels per plane
        vol = torch.rand(D, P, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
        taps = torch.tensor([0.25, 0.5, -0.75], dtype=torch.float64)

        def zconv(padded):                                     # valid 3-tap correlation along z of a shelled tensor
            return taps[0] * padded[:-2] + taps[1] * padded[1:-1] + taps[2] * padded[2:]

        def reflect_shell(t):                                  # [d, P] -> [d + 2, P] with reflect copies
            return torch.cat([t[1:2], t, t[-2:-1]])

        # unsplit reference
        a = zconv(reflect_shell(vol)); a = (a - a.mean()) / a.std(unbiased=False)
        b = zconv(reflect_shell(a)); want = (b - b.mean()) / b.std(unbiased=False)

        lo, hi = (0, D // 2) if rank == 0 else (D // 2, D)
        d = hi - lo
        idx = [lo - 1 if lo > 0 else 1] + list(range(lo, hi)) + [hi if hi < D else D - 2]
        bufs = {0: vol[idx].clone(), 1: torch.zeros(d + 2, P, dtype=torch.float64), 2: torch.zeros(d + 2, P, dtype=torch.float64)}
        sums = {1: torch.zeros(2, dtype=torch.float64), 3: torch.zeros(2, dtype=torch.float64)}
        steps = [(1, 1, 0, 1, "conv"), (4, 1, 0, 1, "norm"), (1, 2, 0, 1, "conv"), (4, 2, 0, 1, "norm")]
        src_of = {0: 0, 2: 1}

        def run_step(i):
            kind, buf = steps[i][0], steps[i][1]
            if kind == 1:                                      # conv: raw output + reflect shell + local sums
                y = zconv(bufs[src_of[i]])
                bufs[buf] = reflect_shell(y)
                sums[i + 1][0], sums[i + 1][1] = y.sum(), (y * y).sum()
            else:                                              # normalise in place, shell included
                s, q = sums[i]
                mean = s / (D * P)
                var = q / (D * P) - mean * mean
                bufs[buf] = (bufs[buf] - mean) / var.sqrt()

        def exchange(buf, goff, groups):
            t = bufs[buf]
            lo_in, hi_in = torch.empty(P, dtype=torch.float64), torch.empty(P, dtype=torch.float64)
            exchange_halo_planes(t[1].contiguous(), t[d].contiguous(), lo_in, hi_in, rank, world)
            if rank > 0:
                t[0] = lo_in
            if rank < world - 1:
                t[d + 1] = hi_in

        run_slab_program(steps, run_step, lambda i: sums[i + 1] if steps[i][0] == 1 else None,
                         lambda t: dist.all_reduce(t), exchange, world)
        assert torch.allclose(bufs[2][1:-1], want[lo:hi], atol=1e-12), (bufs[2][1:-1] - want[lo:hi]).abs().max()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_slab_protocol_with_whole_volume_statistics():
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_slab_worker, args=(2, port), nprocs=2, join=True)


@pytest.mark.parametrize("ragged", [False, True])
def test_two_rank_gloo_sharding_and_halo(ragged):
    port = 29500 + (os.getpid() % 2000) + (1 if ragged else 0)
    mp.spawn(_worker, args=(2, port, ragged), nprocs=2, join=True)
