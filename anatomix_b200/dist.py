"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on B200 boxes; gloo in the CPU tests).

Two ways the path shards:

* **Batch of volumes** — samples are independent (eval-BatchNorm and InstanceNorm
  are per-sample), so ranks take contiguous slices of the batch with no
  data-path collective; an optional all-gather returns every rank's features
  to all ranks (`ShardedExtractor`).
* **One oversized volume** — depth slabs with one-plane halo exchanges after
  every producer step (`slab_bounds`, `HaloSchedule`; the engine side lives in
  `anatomix_b200.halo`).

Nothing here touches the arithmetic: `compute` is any callable mapping a
``[n, C_in, D, H, W]`` tensor to ``[n, C_out, D, H, W]`` (the engine on GPUs, the
oracle in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``n_total`` units for ``rank``: the
    first ``n_total % world`` ranks get one extra unit."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_total: int, world: int) -> List[int]:
    return [shard_range(n_total, world, r)[1] - shard_range(n_total, world, r)[0] for r in range(world)]


class ShardedExtractor:
    """Batch-sharded feature extraction with an optional feature all-gather.

    Every rank calls ``extract(batch)`` with the SAME global batch (or only its
    own slice with ``presharded=True``); the result is this rank's slice, or the
    whole ``[N, C_out, D, H, W]`` tensor on every rank with ``gather=True``.
    """

    def __init__(self, compute: Callable[[torch.Tensor], torch.Tensor], out_channels: int,
                 group: Optional[dist.ProcessGroup] = None):
        self.compute = compute
        self.out_channels = out_channels
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def extract(self, batch: torch.Tensor, gather: bool = False, presharded: bool = False,
                n_total: Optional[int] = None) -> torch.Tensor:
        if presharded:
            if n_total is None:
                raise ValueError("presharded batches need n_total (global batch size)")
            mine = batch
        else:
            n_total = batch.shape[0]
            lo, hi = shard_range(n_total, self.world, self.rank)
            mine = batch[lo:hi]
        out = self.compute(mine) if mine.shape[0] else \
            batch.new_zeros((0, self.out_channels) + tuple(batch.shape[2:]), dtype=torch.float32)
        if not gather or self.world == 1:
            return out
        sizes = shard_sizes(n_total, self.world)
        full = out.new_empty((n_total, self.out_channels) + tuple(out.shape[2:]))
        if len(set(sizes)) == 1:
            # equal shards: one all-gather straight into the output tensor
            dist.all_gather_into_tensor(full, out.contiguous(), group=self.group)
        else:
            # ragged shards: per-rank broadcasts into slices of the output
            off = 0
            for r, s in enumerate(sizes):
                piece = full[off:off + s]
                if r == self.rank:
                    piece.copy_(out)
                if s:
                    dist.broadcast(piece, src=dist.get_global_rank(self.group, r) if self.group else r,
                                   group=self.group)
                off += s
        return full


class FusedGatherExtractor:
    """Batch-sharded extraction where the all-gather is fused into the last conv:
    its epilogue stores every output tile into all ranks' gather buffers over NVLink
    (peer pointers from torch symmetric memory), so no separate collective runs.

    ``extract(batch_shard)`` takes this rank's ``[n, C_in, D, H, W]`` CUDA shard (equal
    ``n`` on every rank) and returns the full ``[world*n, C_out, D, H, W]`` tensor,
    which lives in symmetric memory and is overwritten by the next call."""

    def __init__(self, engine, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        self.engine = engine
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("at most 8 peers (one NVSwitch box)")
        self._buf = None
        self._hdl = None

    def _buffers(self, shape):
        if self._buf is None or tuple(self._buf.shape) != tuple(shape):
            self._buf = self._symm.empty(shape, dtype=torch.float32, device=self.engine.device)
            self._hdl = self._symm.rendezvous(self._buf, self.group)
        return self._buf, self._hdl

    def extract(self, shard: torch.Tensor) -> torch.Tensor:
        n, _, d, h, w = shard.shape
        buf, hdl = self._buffers((self.world * n, self.engine.output_nc, d, h, w))
        hdl.barrier()                                   # peers are done reading the previous result
        self.engine.forward_allgather(shard, list(hdl.buffer_ptrs), self.rank)
        hdl.barrier()                                   # every rank's stores have landed everywhere
        return buf


# ------------------------------------------------------------------ depth slabs
def slab_bounds(depth: int, world: int, num_downs: int) -> List[Tuple[int, int]]:
    """Depth ranges [z_lo, z_hi) per rank for one volume of ``depth`` planes.

    Boundaries are multiples of ``2**num_downs`` so that 2x2x2 pooling and x2
    nearest upsampling never straddle a slab boundary at any level, and every slab
    keeps at least ``2 * 2**num_downs`` planes (the bottleneck needs >= 2 planes for
    reflect padding at the global faces)."""
    unit = 1 << num_downs
    if depth % unit:
        raise ValueError(f"depth {depth} is not a multiple of {unit}")
    blocks = depth // unit
    if blocks < 2 * world:
        raise ValueError(f"depth {depth} is too small for {world} slabs of at least {2 * unit} planes")
    out = []
    for r in range(world):
        lo, hi = shard_range(blocks, world, r)
        out.append((lo * unit, hi * unit))
    return out


def exchange_halo_planes(lo_plane_out: torch.Tensor, hi_plane_out: torch.Tensor,
                         lo_plane_in: torch.Tensor, hi_plane_in: torch.Tensor,
                         rank: int, world: int, group: Optional[dist.ProcessGroup] = None) -> None:
    """One halo step for a tensor split along depth: send my first interior plane to
    rank-1 and my last interior plane to rank+1; receive their boundary planes into
    ``lo_plane_in`` (from rank-1) and ``hi_plane_in`` (from rank+1).  Ranks at the
    global faces skip the missing neighbour (their shell keeps the reflect copy)."""
    # `rank` / `world` are GROUP-local; P2POp addresses peers by GLOBAL rank
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, lo_plane_out, peer(rank - 1), group))
        ops.append(dist.P2POp(dist.irecv, lo_plane_in, peer(rank - 1), group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, hi_plane_out, peer(rank + 1), group))
        ops.append(dist.P2POp(dist.irecv, hi_plane_in, peer(rank + 1), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
