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


class FeatureGather:
    """Batch-sharded extraction with the feature all-gather done WITHOUT a library collective, over peer-mapped
    gather buffers (torch symmetric memory -> NVLink addresses):

    * ``mode="push"`` (default): the forward writes this rank's slice into its own gather buffer, then the copy
      engines push that slice to every peer (`anx_push_to_peers`) on a side stream -- no SM is involved, so
      the pushes of step k overlap the convs of step k + 1 (`submit` / `wait`);
    * ``mode="fused"``: the last conv's epilogue stores every tile into all ranks' buffers itself
      (`anx_engine_forward_gather`); the transfer rides inside the conv kernel.

    ``payload="f32"`` gathers the reference's fp32 NCDHW features (bit-identical to a local forward);
    ``payload="cl16"`` gathers 16-bit channels-last ``[N, D, H, W, C]`` features (half the NVLink bytes; the
    fp32 results rounded once), which `Engine.widen` turns into fp32 NCDHW locally when asked.

    Every rank calls `extract` / `submit` with its own ``[n, C_in, D, H, W]`` CUDA shard (equal ``n``) the same
    number of times.  Results live in symmetric memory and are overwritten ``depth`` submissions later."""

    def __init__(self, engine, group: Optional[dist.ProcessGroup] = None, mode: str = "push", payload: str = "f32",
                 depth: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        if mode not in ("push", "fused") or payload not in ("f32", "cl16"):
            raise ValueError("mode is 'push' or 'fused', payload 'f32' or 'cl16'")
        self._symm = symm_mem
        self.engine, self.mode, self.payload, self.depth = engine, mode, payload, depth
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("at most 8 peers (one NVSwitch box)")
        self._shape = None
        self._bufs, self._hdls, self._done = [], [], []
        self._comm = torch.cuda.Stream(device=engine.device)
        self._k = 0

    def _full_shape(self, n, d, h, w):
        c = self.engine.output_nc
        return (self.world * n, c, d, h, w) if self.payload == "f32" else (self.world * n, d, h, w, c)

    def _ensure(self, n, d, h, w):
        shape = self._full_shape(n, d, h, w)
        if self._shape != shape:
            dtype = torch.float32 if self.payload == "f32" else self.engine.storage_dtype
            self._bufs = [self._symm.empty(shape, dtype=dtype, device=self.engine.device) for _ in range(self.depth)]
            self._hdls = [self._symm.rendezvous(b, self.group) for b in self._bufs]
            self._done = [None] * self.depth
            self._shape = shape

    def submit(self, shard: torch.Tensor) -> int:
        """Queues forward + gather of one shard; returns the slot to pass to `wait`."""
        from . import _lib
        n, _, d, h, w = shard.shape
        self._ensure(n, d, h, w)
        b = self._k % self.depth
        self._k += 1
        buf, hdl = self._bufs[b], self._hdls[b]
        dev = self.engine.device
        main = torch.cuda.current_stream(dev)
        mine = buf[self.rank * n:(self.rank + 1) * n]
        if self._done[b] is not None:
            # the copy engines may still be reading this slot's previous content (push of `depth` submissions ago)
            main.wait_event(self._done[b])
        if self.mode == "fused":
            hdl.barrier()                               # every rank is done reading this buffer's previous content
            self.engine.forward_gather(shard, list(hdl.buffer_ptrs), self.rank,
                                       _lib.PAYLOAD_F32_NCDHW if self.payload == "f32" else _lib.PAYLOAD_CL16)
            hdl.barrier()                               # every rank's stores have landed everywhere
            ev = torch.cuda.Event()
            ev.record(main)
            self._done[b] = ev
            return b
        if self.payload == "f32":
            self.engine.forward(shard, out=mine)
        else:
            self.engine.forward_cl16(shard, out=mine)
        ready = torch.cuda.Event()
        ready.record(main)
        slot_bytes = mine.numel() * mine.element_size()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            hdl.barrier()                               # peers have consumed what this buffer held before
            self.engine.push_to_peers(mine, [p + self.rank * slot_bytes for p in hdl.buffer_ptrs], self.rank)
            hdl.barrier()                               # all slices have landed in every rank's buffer
            ev = torch.cuda.Event()
            ev.record(self._comm)
        self._done[b] = ev
        return b

    def wait(self, slot: int) -> torch.Tensor:
        """The gathered tensor of `slot`; the current stream waits until every rank's slice is in place."""
        torch.cuda.current_stream(self.engine.device).wait_event(self._done[slot])
        return self._bufs[slot]

    def extract(self, shard: torch.Tensor) -> torch.Tensor:
        return self.wait(self.submit(shard))


class FusedGatherExtractor(FeatureGather):
    """`FeatureGather` in its fused form (kept under its round-1 name): the last conv's epilogue stores into all
    ranks' fp32 gather buffers."""

    def __init__(self, engine, group: Optional[dist.ProcessGroup] = None):
        super().__init__(engine, group, mode="fused", payload="f32", depth=1)


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
