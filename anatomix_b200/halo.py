"""Depth-halo spatial partition: ONE oversized volume split along depth over the
GPUs of a box (BASELINE config 5).

Each rank owns a slab of ``D / world`` planes (boundaries multiples of
``2**num_downs`` so pooling and nearest upsampling stay local at every level) and
runs the ordinary engine on it, launch by launch.  After every launch that
produces an activation tensor, neighbouring ranks swap ONE boundary plane of that
tensor: my first interior plane goes into the upper shell plane of rank-1, my
last interior plane into the lower shell plane of rank+1.  The engine's buffers
already carry a one-voxel shell for reflect padding, so a received neighbour
plane simply replaces the mirror copy at an interior slab face and every kernel
runs unchanged; reflect padding survives only at the two global faces.  The
network input is handed over with the neighbour planes attached
(ANX_FLAG_DEPTH_HALO_INPUT).

InstanceNorm networks (`anatomix-dev`): the statistics are per whole volume, so the
sums every conv accumulates (doubles, from its fp32 accumulators) are all-reduced
over the slabs before the normalisation step runs (`anx_engine_step_stats`), and the
normalisation divides by the whole volume's voxel count (`anx_engine_set_slab`).
Trilinear upsampling reads the neighbour's boundary plane from the shell at interior
slab faces.  Average / max pooling and nearest upsampling are slab-local because the
boundaries are multiples of ``2**num_downs``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .dist import exchange_halo_planes, slab_bounds
from .engine import Engine


def run_slab_program(steps, run_step, step_stats, all_reduce_stats, exchange, world: int) -> None:
    """The per-launch protocol of a depth slab, independent of what executes the launches.

    ``steps``: ``[(kind, out_buffer, out_group_offset, out_groups, name)]`` as from `Engine.step_table`.
    After launch i: if it accumulated InstanceNorm sums (``step_stats(i)`` is not None) those are summed over
    the slabs -- the normalisation launch that follows needs whole-volume statistics, and the raw planes need no
    exchange because the normalised ones are swapped after that launch; otherwise, if the launch produced an
    activation tensor, neighbours swap its boundary planes."""
    for i, (kind, buf, goff, groups, name) in enumerate(steps):
        run_step(i)
        if world == 1:
            continue
        stats = step_stats(i)
        if stats is not None:
            all_reduce_stats(stats)
        elif buf >= 0:
            exchange(buf, goff, groups)


def slab_input_with_halo(volume: torch.Tensor, z_lo: int, z_hi: int) -> torch.Tensor:
    """``volume[:, :, z_lo-1 : z_hi+1]`` with reflect copies where the slab touches a
    global face (plane -1 -> plane 1, plane D -> plane D-2)."""
    depth = volume.shape[2]
    idx = [z_lo - 1 if z_lo > 0 else 1] + list(range(z_lo, z_hi)) + [z_hi if z_hi < depth else depth - 2]
    return volume.index_select(2, torch.tensor(idx, device=volume.device)).contiguous()


class DepthSlabExtractor:
    """Feature extraction of one ``[1, C, D, H, W]`` volume with D split over the
    ranks of ``group``.  Every rank passes the same full volume (host or device
    tensor) and gets its own slab of features, or the whole feature volume with
    ``gather=True``."""

    def __init__(self, cfg: dict, state: dict, device, group: Optional[dist.ProcessGroup] = None,
                 in_engine_exchange: Optional[bool] = None):
        """``in_engine_exchange``: run the halo protocol inside the engine (`anx_engine_forward_slab`: peer stores
        into the neighbours' shells + flag words, no host code between launches) instead of the step-wise
        NCCL send/recv protocol.  Default: whenever it applies (several ranks, CUDA, symmetric memory available,
        no InstanceNorm -- whose statistics need an all-reduce between launches --, slabs of equal depth)."""
        if cfg.get("norm", "batch") not in ("batch", "none", "instance") \
                or cfg.get("interp", "nearest") not in ("nearest", "trilinear"):
            raise NotImplementedError("depth-slab mode covers batch / instance / no norm with nearest / trilinear upsampling")
        self.cfg = cfg
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.engine = Engine(cfg, device, flags=_lib.FLAG_DEPTH_HALO_INPUT)
        self.engine.load_state(state)
        self.steps = self.engine.step_table()
        can = self.world > 1 and cfg.get("norm", "batch") != "instance"
        self.in_engine_exchange = can if in_engine_exchange is None else (in_engine_exchange and can)
        self._symm_ws = {}           # (n, d, h, w) -> (workspace, handle, flags, flags handle)

    def _peer_workspace(self, n, d, h, w):
        """This rank's workspace and flag words in symmetric memory (peer-mapped), allocated once per shape."""
        key = (n, d, h, w)
        got = self._symm_ws.get(key)
        if got is None:
            import torch.distributed._symmetric_memory as symm_mem
            grp = self.group if self.group is not None else dist.group.WORLD
            ws = symm_mem.empty(self.engine.workspace_bytes(n, d, h, w), dtype=torch.uint8, device=self.engine.device)
            ws.zero_()
            flags = symm_mem.empty(64, dtype=torch.int32, device=self.engine.device)
            flags.zero_()
            h_ws, h_fl = symm_mem.rendezvous(ws, grp), symm_mem.rendezvous(flags, grp)
            torch.cuda.synchronize(self.engine.device)
            h_fl.barrier()                      # every rank's flags are zero before anyone publishes
            got = (ws, h_ws, flags, h_fl)
            self._symm_ws = {key: got}          # one shape at a time: these buffers are large
        return got

    def forward_slab(self, x: torch.Tensor, out: torch.Tensor):
        """One depth-slab forward through `anx_engine_forward_slab` on this rank's slab input `x` (neighbour
        planes attached) into `out`; asynchronous on the current stream."""
        n, d, h, w = out.shape[0], out.shape[2], out.shape[3], out.shape[4]
        ws, h_ws, flags, h_fl = self._peer_workspace(n, d, h, w)
        lo, hi = self.rank - 1, self.rank + 1
        self.engine.forward_slab(
            x, out, ws,
            h_ws.buffer_ptrs[lo] if lo >= 0 else 0, h_ws.buffer_ptrs[hi] if hi < self.world else 0,
            h_fl.buffer_ptrs[self.rank],
            h_fl.buffer_ptrs[lo] if lo >= 0 else 0, h_fl.buffer_ptrs[hi] if hi < self.world else 0)

    def _exchange(self, ws, table, buf, goff, groups, n, d, h, w):
        off, nbytes, level, gtot = table[buf]
        dl, hl, wl = d >> level, h >> level, w >> level
        plane = (hl + 2) * self.engine.row_layout(wl)[1] * 16      # whole padded rows, lead / tail voxels included
        view = ws[off:off + n * gtot * (dl + 2) * plane].view(n, gtot, dl + 2, plane)[:, goff:goff + groups]
        lo_out, hi_out = view[:, :, 1].contiguous(), view[:, :, dl].contiguous()
        lo_in, hi_in = torch.empty_like(lo_out), torch.empty_like(hi_out)
        exchange_halo_planes(lo_out, hi_out, lo_in, hi_in, self.rank, self.world, self.group)
        if self.rank > 0:
            view[:, :, 0] = lo_in
        if self.rank < self.world - 1:
            view[:, :, dl + 1] = hi_in

    @torch.no_grad()
    def extract(self, volume: torch.Tensor, gather: bool = False) -> torch.Tensor:
        n, _, depth, h, w = volume.shape
        z_lo, z_hi = slab_bounds(depth, self.world, self.cfg["num_downs"])[self.rank]
        d = z_hi - z_lo
        self.engine.set_slab(self.rank > 0, self.rank < self.world - 1, depth if self.world > 1 else 0)
        x = slab_input_with_halo(volume, z_lo, z_hi).to(self.engine.device, torch.float32)
        out = torch.empty((n, self.cfg["output_nc"], d, h, w), dtype=torch.float32, device=self.engine.device)
        bounds = slab_bounds(depth, self.world, self.cfg["num_downs"])
        if self.in_engine_exchange and len({b[1] - b[0] for b in bounds}) == 1:
            self.forward_slab(x, out)
            return self._gather(out, depth) if gather else out
        ws = self.engine.workspace(n, d, h, w)
        table = self.engine.buffer_table(n, d, h, w)
        def stats_view(i):
            off, nbytes = self.engine.step_stats(i, n, d, h, w)
            return ws[off:off + nbytes].view(torch.float64) if nbytes else None

        run_slab_program(
            self.steps,
            run_step=lambda i: self.engine.run_steps(x, out, i, i + 1),
            step_stats=stats_view,
            all_reduce_stats=lambda t: dist.all_reduce(t, group=self.group),
            exchange=lambda buf, goff, groups: self._exchange(ws, table, buf, goff, groups, n, d, h, w),
            world=self.world)
        if not gather or self.world == 1:
            return out
        return self._gather(out, depth)

    def _gather(self, out: torch.Tensor, depth: int) -> torch.Tensor:
        # slabs have equal depth unless depth/unit is not a multiple of world: gather plane-major, then permute back
        sizes = [b[1] - b[0] for b in slab_bounds(depth, self.world, self.cfg["num_downs"])]
        if len(set(sizes)) != 1:
            raise NotImplementedError("gather with unequal slabs")
        pieces = [torch.empty_like(out) for _ in range(self.world)]
        dist.all_gather(pieces, out, group=self.group)
        return torch.cat(pieces, dim=2)
