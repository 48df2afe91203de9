"""Python host side of the B200 engine: an `Engine` object over the C ABI
(``include/anatomix_b200.h``) plus the glue that lets the `Unet` module hand
eligible forwards to it.

torch is used here only as plumbing: device memory (workspace / output
tensors from the caching allocator), the current CUDA stream, and reading the
module's parameters.  All arithmetic of an eligible forward happens inside
``libanatomix_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import operator
import os
import weakref
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import EngineError


def _desc_from_cfg(cfg: dict, device_index: int, flags: int = 0) -> _lib.UnetDesc:
    d = _lib.UnetDesc()
    d.struct_size = C.sizeof(_lib.UnetDesc)
    d.input_nc, d.output_nc = cfg["input_nc"], cfg["output_nc"]
    d.num_downs, d.ngf = cfg["num_downs"], cfg.get("ngf", 24)
    d.norm_kind = _lib.NORM[cfg.get("norm", "batch")]
    d.norm_eps = cfg.get("norm_eps", 1e-5)
    d.act_kind = _lib.ACT[cfg.get("activation", "relu")]
    d.act_slope = 0.3                      # reference network.py:191
    d.pool_kind = _lib.POOL[cfg.get("pooling", "Max")]
    d.interp_kind = _lib.INTERP[cfg.get("interp", "nearest")]
    d.device = device_index
    d.flags = flags
    return d


class Engine:
    """One network on one CUDA device.  ``cfg`` holds the reference constructor
    kwargs (network.py:262-279); parameters arrive via `load_state`."""

    def __init__(self, cfg: dict, device: torch.device | int | str = "cuda", flags: int = 0):
        self.lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("the engine runs on CUDA devices only")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.cfg = dict(cfg)
        self._h = C.c_void_p()
        desc = _desc_from_cfg(cfg, self.device.index, flags)
        with torch.cuda.device(self.device):
            st = self.lib.anx_engine_create(C.byref(desc), C.byref(self._h))
        if st != _lib.ANX_OK:
            raise EngineError(st, self.lib.anx_status_string(st).decode())
        self.output_nc = cfg["output_nc"]      # channels a forward writes (a fused head changes it)
        self.input_nc = cfg["input_nc"]
        self.flags = flags
        self._workspaces: Dict[tuple, torch.Tensor] = {}
        self._state, self._tap_convs_ready = None, set()

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.anx_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != _lib.ANX_OK:
            raise EngineError(st, self.lib.anx_engine_last_error(self._h).decode())

    # -- parameters ----------------------------------------------------------
    def conv_table(self):
        """[(module_index, cin, cout, has_norm)] in network order."""
        out = []
        for k in range(self.lib.anx_engine_num_convs(self._h)):
            v = [C.c_int32() for _ in range(4)]
            self._check(self.lib.anx_engine_conv_info(self._h, k, *[C.byref(x) for x in v]))
            out.append(tuple(x.value for x in v))
        return out

    def load_state(self, state: dict):
        """Feeds a reference-format state dict (``model.<idx>.weight`` ...)."""
        norm = self.cfg.get("norm", "batch")
        self._state = state            # pre-norm tap clones are packed from it on first use (`_ensure_tap_conv`)
        self._tap_convs_ready = set()
        for k, (idx, cin, cout, has_norm) in enumerate(self.conv_table()):
            def host(name):
                t = torch.as_tensor(state[name]).detach().to("cpu", torch.float32).contiguous()
                return t
            w = host(f"model.{idx}.weight")
            if tuple(w.shape) != (cout, cin, 3, 3, 3):
                raise ValueError(f"model.{idx}.weight has shape {tuple(w.shape)}, expected {(cout, cin, 3, 3, 3)}")
            keep = [w]
            b = None
            if f"model.{idx}.bias" in state:
                b = host(f"model.{idx}.bias"); keep.append(b)
            bn = [None] * 4
            if has_norm and norm == "batch":
                bn = [host(f"model.{idx + 1}.{n}") for n in ("weight", "bias", "running_mean", "running_var")]
                keep += bn
            ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            with torch.cuda.device(self.device):
                self._check(self.lib.anx_engine_set_conv(self._h, k, ptr(w), ptr(b), *[ptr(t) for t in bn], 0))

    def set_head(self, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor] = None):
        """Fuses a 1x1x1 conv ``[K, output_nc(,1,1,1)]`` (+ bias) into the last conv's epilogue
        (anx_engine_set_head); ``None`` removes it.  Forwards then return ``[N, K, D, H, W]``."""
        if weight is None:
            self._check(self.lib.anx_engine_set_head(self._h, 0, None, None, 0))
        else:
            w = weight.detach().to("cpu", torch.float32).reshape(weight.shape[0], -1).contiguous()
            if w.shape[1] != self.cfg["output_nc"]:
                raise ValueError(f"head weight has {w.shape[1]} input channels, the network outputs {self.cfg['output_nc']}")
            b = None if bias is None else bias.detach().to("cpu", torch.float32).contiguous()
            with torch.cuda.device(self.device):
                self._check(self.lib.anx_engine_set_head(
                    self._h, w.shape[0], C.c_void_p(w.data_ptr()), None if b is None else C.c_void_p(b.data_ptr()), 0))
        self.output_nc = self.lib.anx_engine_out_channels(self._h)

    # -- feature taps ------------------------------------------------------------
    def tap_table(self):
        """{module_index: (ordinal, channels, level, last_step, is_output, prenorm_conv)} of the tensors the
        engine can hand out for ``forward(layers=[...])`` (anx_engine_tap_info / anx_engine_tap_kind).
        ``prenorm_conv`` >= 0: the PRE-norm output of that conv ordinal, re-evaluated by an un-folded clone."""
        out = {}
        for k in range(self.lib.anx_engine_num_taps(self._h)):
            v = [C.c_int32() for _ in range(5)]
            self._check(self.lib.anx_engine_tap_info(self._h, k, *[C.byref(x) for x in v]))
            kind, conv = C.c_int32(), C.c_int32()
            self._check(self.lib.anx_engine_tap_kind(self._h, k, C.byref(kind), C.byref(conv)))
            out[v[0].value] = (k, v[1].value, v[2].value, v[3].value, bool(v[4].value),
                               conv.value if kind.value == 2 else -1)
        return out

    def _ensure_tap_conv(self, conv_ordinal: int):
        """Packs the un-folded clone of conv `conv_ordinal` (plain weight + bias) for pre-norm taps."""
        if conv_ordinal in self._tap_convs_ready:
            return
        idx = self.conv_table()[conv_ordinal][0]
        w = torch.as_tensor(self._state[f"model.{idx}.weight"]).detach().to("cpu", torch.float32).contiguous()
        b = self._state.get(f"model.{idx}.bias")
        b = None if b is None else torch.as_tensor(b).detach().to("cpu", torch.float32).contiguous()
        with torch.cuda.device(self.device):
            self._check(self.lib.anx_engine_set_tap_conv(
                self._h, conv_ordinal, C.c_void_p(w.data_ptr()), None if b is None else C.c_void_p(b.data_ptr()), 0))
        self._tap_convs_ready.add(conv_ordinal)

    def forward_taps(self, x: torch.Tensor, layers, encode_only: bool = False):
        """``Unet.forward(x, layers, encode_only)`` of the reference (network.py:475-529) for tap
        indices the engine materialises: ``(output, taps)``, or ``taps`` alone when ``encode_only``
        stops at ``layers[-1]``.  Taps come back in slot order, fp32 NCDHW."""
        table = self.tap_table()
        x = x.contiguous().float()
        n, _, d, h, w = x.shape
        ws = self.workspace(n, d, h, w)
        steps = self.lib.anx_engine_num_steps(self._h)
        stop_early = encode_only and layers[-1] in table
        wanted = sorted({i for i in layers if i in table and (not stop_early or i <= layers[-1])})
        last = table[layers[-1]][3] + 1 if stop_early else steps
        out = torch.empty((n, self.output_nc, d, h, w), dtype=torch.float32, device=self.device)
        taps = {}
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            # The workspace shares memory between tensors whose lifetimes do not overlap, so a tapped tensor is
            # exported right after the launch that completes it, before later launches may reuse its region.
            done = 0
            for idx in sorted(wanted, key=lambda i: table[i][3]):
                k, ch, lvl, last_step, is_output, prenorm_conv = table[idx]
                upto = min(last_step + 1, last)
                if upto > done:
                    self._check(self.lib.anx_engine_run_steps(
                        self._h, x.data_ptr(), out.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), stream, done, upto))
                    done = upto
                if is_output:
                    taps[idx] = out
                    continue
                t = torch.empty((n, ch, d >> lvl, h >> lvl, w >> lvl), dtype=torch.float32, device=self.device)
                if prenorm_conv >= 0:       # conv(x) + bias before the norm: the conv's launch again, un-folded
                    self._ensure_tap_conv(prenorm_conv)
                    self._check(self.lib.anx_engine_export_prenorm_tap(
                        self._h, k, x.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), t.data_ptr(), stream))
                else:
                    self._check(self.lib.anx_engine_export_tap(
                        self._h, k, n, d, h, w, ws.data_ptr(), ws.numel(), t.data_ptr(), stream))
                taps[idx] = t
            if last > done:
                self._check(self.lib.anx_engine_run_steps(
                    self._h, x.data_ptr(), out.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), stream, done, last))
        taps = [taps[i] for i in wanted]
        return taps if stop_early else (out, taps)

    # -- forward ---------------------------------------------------------------
    def workspace_bytes(self, n, d, h, w) -> int:
        return self.lib.anx_engine_workspace_bytes(self._h, n, d, h, w)

    def launches_per_forward(self, n, d, h, w) -> int:
        return self.lib.anx_engine_launches_per_forward(self._h, n, d, h, w)

    def _shape_error(self, shape):
        unit = 1 << self.cfg["num_downs"]
        return ValueError(
            f"input of shape {tuple(shape)} is not usable by this U-Net: each of D, H, W must be a "
            f"multiple of {unit} and at least {2 * unit} (the reference fails on such shapes too)")

    def workspace(self, n, d, h, w) -> torch.Tensor:
        """Activation workspace for one shape ON THE CURRENT STREAM: forwards of the same shape queued on
        different streams (DataParallel threads, user side streams) get distinct buffers, so they never race
        on the activations (include/anatomix_b200.h: concurrent forwards need distinct workspaces)."""
        stream = torch.cuda.current_stream(self.device)
        key = (stream.cuda_stream, n, d, h, w)
        ws = self._workspaces.get(key)
        if ws is None:
            need = self.workspace_bytes(n, d, h, w)
            if need == 0:
                raise self._shape_error((n, self.input_nc, d, h, w))
            if len(self._workspaces) >= 4:
                old_key = next(iter(self._workspaces))
                old = self._workspaces.pop(old_key)
                # work queued on the evicted workspace's stream may still read it: tell the caching
                # allocator not to hand the block out before that stream has passed this point
                old.record_stream(torch.cuda.ExternalStream(old_key[0], device=self.device)
                                  if old_key[0] else torch.cuda.default_stream(self.device))
                del old
            # zeroed once: the unused lead / tail voxels of every padded row then stay finite and equal
            # between runs (halo plane exchanges and buffer dumps copy them along)
            with torch.cuda.device(self.device):
                ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
            self._workspaces[key] = ws
        return ws

    def _check_out(self, out: torch.Tensor, shape, what="out"):
        if not isinstance(out, torch.Tensor) or tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 \
                or out.device != self.device or not out.is_contiguous():
            raise ValueError(f"`{what}` must be a contiguous fp32 tensor of shape {tuple(shape)} on {self.device}; "
                             f"got {tuple(out.shape)} {out.dtype} on {out.device}"
                             f"{'' if out.is_contiguous() else ' (non-contiguous)'}")

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 NCDHW CUDA tensor in, fp32 NCDHW CUDA tensor out, queued on the
        current stream of the engine's device."""
        if x.dim() != 5 or x.shape[1] != self.input_nc:
            raise ValueError(f"expected input [N, {self.input_nc}, D, H, W], got {tuple(x.shape)}")
        if x.device != self.device:
            raise ValueError(f"input on {x.device}, engine on {self.device}")
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        n, _, d, h, w = x.shape
        ws = self.workspace(n, d, h, w)
        if out is None:
            out = torch.empty((n, self.output_nc, d, h, w), dtype=torch.float32, device=self.device)
        else:
            self._check_out(out, (n, self.output_nc, d, h, w))
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward(
                self._h, x.data_ptr(), out.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), stream))
        return out

    def forward_into(self, x: torch.Tensor, dest: torch.Tensor, channel_offset: int) -> torch.Tensor:
        """Forward whose output lands in channels ``[channel_offset, channel_offset + C)`` of the wider contiguous
        fp32 tensor ``dest`` ``[N, C_total, D, H, W]`` -- a zero-copy ``torch.cat`` with features the caller puts into
        the other channels (anx_engine_forward_concat; reference instance_optimization.py:16-119 concatenates the
        MIND-SSC descriptors in front of the network features)."""
        x = x.contiguous().float()
        n, _, d, h, w = x.shape
        if dest.dim() != 5 or dest.shape[0] != n or tuple(dest.shape[2:]) != (d, h, w) or dest.dtype != torch.float32 \
                or dest.device != self.device or not dest.is_contiguous() \
                or not 0 <= channel_offset <= dest.shape[1] - self.output_nc:
            raise ValueError(f"`dest` must be a contiguous fp32 [N, >= {channel_offset + self.output_nc}, D, H, W] tensor "
                             f"on {self.device} matching the input's batch and spatial size")
        ws = self.workspace(n, d, h, w)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_concat(
                self._h, x.data_ptr(), dest.data_ptr(), dest.shape[1], channel_offset, n, d, h, w,
                ws.data_ptr(), ws.numel(), stream))
        return dest

    def forward_allgather(self, x: torch.Tensor, peer_ptrs, rank: int):
        """Forward whose last conv stores into every rank's gather buffer (see
        anx_engine_forward_allgather).  ``peer_ptrs``: device pointers (ints) of the
        ranks' [world*n, C, D, H, W] fp32 buffers, NVLink-mapped into this process."""
        x = x.contiguous().float()
        n, _, d, h, w = x.shape
        ws = self.workspace(n, d, h, w)
        world = len(peer_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_allgather(
                self._h, x.data_ptr(), arr, world, rank, n, d, h, w, ws.data_ptr(), ws.numel(), stream))

    # -- 16-bit channels-last payload ---------------------------------------------
    @property
    def storage_dtype(self) -> torch.dtype:
        """torch dtype of stored activations and of the CL16 payload (anx_engine_storage_type)."""
        return torch.bfloat16 if self.lib.anx_engine_storage_type(self._h) == 0 else torch.float16

    def forward_cl16(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Forward that writes the features as a 16-bit channels-last ``[N, D, H, W, C]`` tensor (the fp32
        results rounded once; half the bytes of the drop-in output).  ``out.permute(0, 4, 1, 2, 3)`` is the
        logical NCDHW view (torch's channels_last_3d)."""
        x = x.contiguous().float()
        n, _, d, h, w = x.shape
        ws = self.workspace(n, d, h, w)
        if out is None:
            out = torch.empty((n, d, h, w, self.output_nc), dtype=self.storage_dtype, device=self.device)
        elif tuple(out.shape) != (n, d, h, w, self.output_nc) or out.dtype != self.storage_dtype \
                or out.device != self.device or not out.is_contiguous():
            raise ValueError(f"`out` must be a contiguous {self.storage_dtype} tensor [N, D, H, W, C] on {self.device}")
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_cl16(
                self._h, x.data_ptr(), out.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), stream))
        return out

    def widen(self, cl16: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """16-bit channels-last ``[N, D, H, W, C]`` -> fp32 NCDHW (anx_widen_cl16_f32)."""
        n, d, h, w, c = cl16.shape
        if out is None:
            out = torch.empty((n, c, d, h, w), dtype=torch.float32, device=cl16.device)
        with torch.cuda.device(cl16.device):
            st = self.lib.anx_widen_cl16_f32(cl16.data_ptr(), out.data_ptr(), n, c, d, h, w,
                                             0 if cl16.dtype == torch.bfloat16 else 1,
                                             torch.cuda.current_stream(cl16.device).cuda_stream)
        if st != _lib.ANX_OK:
            raise EngineError(st, "anx_widen_cl16_f32")
        return out

    def forward_gather(self, x: torch.Tensor, peer_ptrs, rank: int, payload: int = _lib.PAYLOAD_F32_NCDHW):
        """Forward whose last conv stores into every rank's gather buffer in the chosen payload
        (anx_engine_forward_gather)."""
        x = x.contiguous().float()
        n, _, d, h, w = x.shape
        ws = self.workspace(n, d, h, w)
        world = len(peer_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_gather(
                self._h, x.data_ptr(), arr, world, rank, payload, n, d, h, w, ws.data_ptr(), ws.numel(), stream))

    def push_to_peers(self, src: torch.Tensor, peer_dst_ptrs, rank: int):
        """Copy-engine push of ``src`` (contiguous) to ``peer_dst_ptrs[r]`` for every r != rank, on the current
        stream (anx_push_to_peers)."""
        world = len(peer_dst_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_dst_ptrs])
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_push_to_peers(self._h, src.data_ptr(), arr, world, rank,
                                                   src.numel() * src.element_size(), stream))

    def forward_slab(self, x: torch.Tensor, out: torch.Tensor, ws: torch.Tensor, lower_ws: int, upper_ws: int,
                     flags: int, lower_flags: int, upper_flags: int):
        """Depth-slab forward with the halo exchange inside the engine (anx_engine_forward_slab).  ``ws`` is this
        rank's peer-visible workspace; the integer arguments are device addresses (0 = no neighbour)."""
        n, d, h, w = out.shape[0], out.shape[2], out.shape[3], out.shape[4]
        links = _lib.SlabLinks(C.sizeof(_lib.SlabLinks), lower_ws or None, upper_ws or None, flags,
                               lower_flags or None, upper_flags or None)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_slab(
                self._h, x.data_ptr(), out.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), C.byref(links), stream))

    def forward_host_cl16(self, x_host: torch.Tensor, out_host: torch.Tensor, dev_in: torch.Tensor,
                          dev_out: torch.Tensor):
        """`forward_host` with the 16-bit channels-last payload: ``out_host`` / ``dev_out`` are
        ``[N, D, H, W, C]`` tensors of `storage_dtype` (half the download of the fp32 drop-in output)."""
        n, _, d, h, w = x_host.shape
        shape = (n, d, h, w, self.output_nc)
        for t, dev, what in ((out_host, "cpu", "out_host"), (dev_out, "cuda", "dev_out")):
            if tuple(t.shape) != shape or t.dtype != self.storage_dtype or t.device.type != dev or not t.is_contiguous():
                raise ValueError(f"`{what}` must be a contiguous {self.storage_dtype} {dev} tensor of shape {shape}")
        self._check_out(dev_in, (n, self.input_nc, d, h, w), "dev_in")
        ws = self.workspace(n, d, h, w)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_host_ex(
                self._h, x_host.data_ptr(), out_host.data_ptr(), _lib.PAYLOAD_CL16, n, d, h, w, dev_in.data_ptr(),
                dev_out.data_ptr(), ws.data_ptr(), ws.numel(), stream))

    def forward_host_pipelined(self, x_host: torch.Tensor, out_host: torch.Tensor, dev_in: torch.Tensor,
                               dev_out: torch.Tensor):
        """`forward_host` for back-to-back calls: the current stream does not wait for the download, so the next
        call's upload and convs overlap it (anx_engine_forward_host_pipelined).  Alternate two ``(dev_out, out_host)``
        sets; results are complete after `host_wait()` + a stream synchronisation."""
        n, _, d, h, w = x_host.shape
        for t, shape, what in ((x_host, (n, self.input_nc, d, h, w), "x_host"),
                               (out_host, (n, self.output_nc, d, h, w), "out_host")):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise ValueError(f"`{what}` must be a contiguous fp32 CPU tensor of shape {shape}")
        self._check_out(dev_in, (n, self.input_nc, d, h, w), "dev_in")
        self._check_out(dev_out, (n, self.output_nc, d, h, w), "dev_out")
        ws = self.workspace(n, d, h, w)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_host_pipelined(
                self._h, x_host.data_ptr(), out_host.data_ptr(), _lib.PAYLOAD_F32_NCDHW, n, d, h, w,
                dev_in.data_ptr(), dev_out.data_ptr(), ws.data_ptr(), ws.numel(), stream))

    def host_wait(self):
        """The current stream waits for every download issued by `forward_host_pipelined` so far."""
        with torch.cuda.device(self.device):
            self._check(self.lib.anx_engine_host_wait(self._h, torch.cuda.current_stream(self.device).cuda_stream))

    def forward_host(self, x_host: torch.Tensor, out_host: torch.Tensor, dev_in: torch.Tensor,
                     dev_out: torch.Tensor):
        """End-to-end call on (pinned) host buffers; see anx_engine_forward_host."""
        if x_host.dim() != 5 or x_host.shape[1] != self.input_nc:
            raise ValueError(f"expected host input [N, {self.input_nc}, D, H, W], got {tuple(x_host.shape)}")
        n, _, d, h, w = x_host.shape
        for t, shape, what in ((x_host, (n, self.input_nc, d, h, w), "x_host"),
                               (out_host, (n, self.output_nc, d, h, w), "out_host")):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise ValueError(f"`{what}` must be a contiguous fp32 CPU tensor of shape {shape}")
        self._check_out(dev_in, (n, self.input_nc, d, h, w), "dev_in")
        self._check_out(dev_out, (n, self.output_nc, d, h, w), "dev_out")
        ws = self.workspace(n, d, h, w)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_forward_host(
                self._h, x_host.data_ptr(), out_host.data_ptr(), n, d, h, w, dev_in.data_ptr(),
                dev_out.data_ptr(), ws.data_ptr(), ws.numel(), stream))

    def profile(self, x: torch.Tensor):
        """[(step name, milliseconds)] of one forward, CUDA-event timed per launch."""
        x = x.contiguous().float()
        n, _, d, h, w = x.shape
        ws = self.workspace(n, d, h, w)
        out = torch.empty((n, self.output_nc, d, h, w), dtype=torch.float32, device=self.device)
        cap = 256
        ms = (C.c_float * cap)()
        names = C.create_string_buffer(32 * cap)
        cnt = C.c_int32()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_profile(
                self._h, x.data_ptr(), out.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), stream,
                ms, names, cap, C.byref(cnt)))
        raw = names.raw
        return [(raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), ms[i]) for i in range(cnt.value)]

    def step_table(self):
        """[(kind, out_buffer, out_group_offset, out_groups, name)] per launch of a forward."""
        out = []
        for i in range(self.lib.anx_engine_num_steps(self._h)):
            v = [C.c_int32() for _ in range(4)]
            name = C.create_string_buffer(32)
            self._check(self.lib.anx_engine_step_info(self._h, i, *[C.byref(x) for x in v], name))
            out.append(tuple(x.value for x in v) + (name.value.decode(),))
        return out

    def run_steps(self, x: torch.Tensor, out: torch.Tensor, first: int, last: int):
        """Runs launches [first, last) of the forward program (see anx_engine_run_steps).
        ``x`` / ``out`` are the same tensors for every call of one forward."""
        n, d, h, w = out.shape[0], out.shape[2], out.shape[3], out.shape[4]
        ws = self.workspace(n, d, h, w)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self._check(self.lib.anx_engine_run_steps(
                self._h, x.data_ptr(), out.data_ptr(), n, d, h, w, ws.data_ptr(), ws.numel(), stream, first, last))

    def row_layout(self, w: int):
        """(lead, pitch) of a padded planar row of interior width ``w`` (anx_engine_row_layout)."""
        lead, pitch = C.c_int32(), C.c_int32()
        self._check(self.lib.anx_engine_row_layout(self._h, w, C.byref(lead), C.byref(pitch)))
        return lead.value, pitch.value

    def set_slab(self, has_lower: bool, has_upper: bool, depth_total: int):
        """Depth-slab mode (anx_engine_set_slab): neighbours at the z faces, depth of the whole volume."""
        self._check(self.lib.anx_engine_set_slab(self._h, int(has_lower), int(has_upper), depth_total))

    def step_stats(self, step: int, n, d, h, w):
        """(offset, bytes) of the InstanceNorm sums the given step accumulates (bytes == 0: none)."""
        off, nb = C.c_size_t(), C.c_size_t()
        self._check(self.lib.anx_engine_step_stats(self._h, step, n, d, h, w, C.byref(off), C.byref(nb)))
        return off.value, nb.value

    def buffer_table(self, n, d, h, w):
        """[(offset, bytes, level, groups)] of the workspace's activation buffers."""
        out = []
        for i in range(self.lib.anx_engine_num_buffers(self._h)):
            off, nb, lvl, grp = C.c_size_t(), C.c_size_t(), C.c_int32(), C.c_int32()
            self._check(self.lib.anx_engine_buffer_info(self._h, n, d, h, w, i, C.byref(off), C.byref(nb),
                                                        C.byref(lvl), C.byref(grp)))
            out.append((off.value, nb.value, lvl.value, grp.value))
        return out


# --------------------------------------------------------------------- eligibility
def _has_hooks(module: nn.Module) -> bool:
    """True when any submodule carries forward hooks / pre-hooks: the engine never calls the submodules,
    so such a model must run the stock loop for the hooks to fire."""
    for m in module.modules():
        if m._forward_hooks or m._forward_pre_hooks:
            return True
    return False


def module_ineligible_reason(module: nn.Module, cfg: dict, device: torch.device) -> Optional[str]:
    """The module-side half of the eligibility test (configuration, modes, parameters, hooks)."""
    if cfg["pad_type"] != "reflect" or not cfg["doubleconv"] or not cfg["use_skip_connection"] \
            or cfg["residual_connection"] or cfg["final_act"] != "none":
        return "non-released topology flags"
    if cfg["norm"] == "batch":
        # the mode of every BatchNorm submodule counts, not only the top-level flag: `model.model[1].train()`
        # on an eval model makes that layer use batch statistics in the reference
        if module.training or any(m.training for m in module.modules()
                                  if isinstance(m, nn.modules.batchnorm._BatchNorm)):
            return "BatchNorm in train mode uses batch statistics"
    elif cfg["norm"] not in ("none", "instance"):
        return f"norm {cfg['norm']!r} is not handled by the engine"
    if cfg["activation"] not in ("relu", "lrelu", "none"):
        return "activation not supported"
    if cfg["pooling"] not in ("Max", "Avg") or cfg["interp"] not in ("nearest", "trilinear"):
        return "pooling / interpolation not supported"
    widths = [cfg["ngf"] << i for i in range(cfg["num_downs"] + 1)]
    if cfg["ngf"] % 16 != 0 or cfg["ngf"] > 64 or widths[-1] > 1024 or any(w > 256 and w % 256 for w in widths) \
            or cfg["input_nc"] > 4 or cfg["output_nc"] > 256:
        return "channel widths outside the tensor-core kernel's range"
    for t in list(module.parameters()) + list(module.buffers()):
        if t.is_floating_point() and (t.dtype != torch.float32 or t.device != device):
            return "parameters are not fp32 on the input's device (the reference raises here)"
    if _has_hooks(module):
        return "forward hooks are registered on the module or a submodule"
    return None


def ineligible_reason(module: nn.Module, cfg: dict, x, layers=()) -> Optional[str]:
    """Why a call must stay on the stock torch path, or None if the engine takes
    it."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        return "input is not a CUDA tensor"
    if x.numel() == 0:
        return "empty batch"
    if cfg["dimension"] != 3 or x.dim() != 5:
        return "not a 3-D network / 5-D input"
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters())):
        return "autograd is recording"
    if torch.is_autocast_enabled():
        return "autocast region"
    why = module_ineligible_reason(module, cfg, x.device)
    if why is not None:
        return why
    if x.shape[1] != cfg["input_nc"]:
        return "channel mismatch"
    unit = 1 << cfg["num_downs"]
    if any(s % unit != 0 or s < 2 * unit for s in x.shape[2:]):
        return "spatial size not a multiple of 2^num_downs (the reference fails here too)"
    if x.dtype != torch.float32:
        return "input is not fp32"
    if len(layers) > 0:
        # feature taps (network.py:475-529): served when every tapped slot is a tensor the engine stores
        missing = binding_for(module, cfg).untappable(x.device, layers)
        if missing:
            return f"feature tap at slot {missing[0]} is not a tensor the engine can hand out"
    return None


# ------------------------------------------------------------------------ bindings
# Engine bindings live OUTSIDE the modules they serve (weakly keyed by the module), so a model that has run on
# the engine still deep-copies, pickles and `torch.save`s like the reference's: ctypes handles never enter a
# module's `__dict__`.  A copy / unpickled module simply gets its own binding on its first eligible forward.
_BINDINGS: "weakref.WeakKeyDictionary[nn.Module, ModuleBinding]" = weakref.WeakKeyDictionary()

# In-place updates through `.data` (``p.data.copy_()``, ``init.normal_(m.weight.data)``: reference
# pretraining_networks.py:695-713) bump neither `_version` nor the storage pointer, so the cheap stamp cannot see
# them.  Every VERIFY_EVERY-th forward therefore also compares a device-side checksum of all parameters and
# buffers (one fused reduction + one 8-byte readback); 1 = every forward.  `invalidate()` forces a re-pack.
VERIFY_EVERY = max(1, int(os.environ.get("ANATOMIX_B200_VERIFY_EVERY", "32")))


def tensors_stamp(tensors) -> tuple:
    return tuple((t.data_ptr(), t._version) for t in tensors)


def tensors_checksum(tensors) -> tuple:
    """Content fingerprint of a list of tensors: (sum, sum of squares weighted by position) as python floats;
    one sync.  Catches `.data` edits the version counters miss."""
    fl = [t.detach().reshape(-1).double() for t in tensors if t.is_floating_point() and t.numel()]
    if not fl:
        return (0.0, 0.0)
    sums = torch.stack([f.sum() for f in fl])
    sq = torch.stack([(f * f).sum() for f in fl])
    k = torch.arange(1, len(fl) + 1, dtype=torch.float64, device=sums.device)
    both = torch.stack([(sums * k).sum(), (sq * k).sum()]).cpu()
    return (both[0].item(), both[1].item())


class PackStamp:
    """When do packed weights need a refresh?  Cheap test every call (storage pointers + version counters),
    content checksum every `VERIFY_EVERY` calls and whenever `invalidate()` was called."""

    def __init__(self):
        self.stamp, self.checksum, self.calls, self.dirty = None, None, 0, True

    def invalidate(self):
        self.dirty = True

    def needs_repack(self, tensors) -> bool:
        st = tensors_stamp(tensors)
        self.calls += 1
        if self.dirty or st != self.stamp:
            self.stamp, self.checksum, self.dirty = st, tensors_checksum(tensors), False
            return True
        if self.calls % VERIFY_EVERY == 0:
            cs = tensors_checksum(tensors)
            if cs != self.checksum:
                self.checksum = cs
                return True
        return False


class ModuleBinding:
    """Keeps one `Engine` per device in sync with a live ``nn.Module``: packed
    weights are rebuilt whenever a parameter or buffer changed (``_version`` bump
    from ``load_state_dict`` / an optimizer step, new storage from ``.to()``, or -- checked every
    `VERIFY_EVERY` forwards -- a content change made through ``.data``)."""

    def __init__(self, module: nn.Module, cfg: dict):
        self._module = weakref.ref(module)
        self.cfg = cfg
        self.engines: Dict[tuple, Engine] = {}
        self.stamps: Dict[tuple, PackStamp] = {}

    @property
    def module(self) -> nn.Module:
        m = self._module()
        if m is None:
            raise RuntimeError("the module of this binding has been collected")
        return m

    def invalidate(self):
        """Forces the next forward to re-pack the weights (after edits the stamp cannot see)."""
        for s in self.stamps.values():
            s.invalidate()

    def engine_for(self, device: torch.device, flags: int = 0) -> Engine:
        key = (device, flags)
        eng = self.engines.get(key)
        if eng is None:
            eng = Engine(self.cfg, device, flags)
            self.engines[key] = eng
            self.stamps[key] = PackStamp()
        m = self.module
        if self.stamps[key].needs_repack(list(m.parameters()) + list(m.buffers())):
            eng.load_state(m.state_dict())
        return eng

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.engine_for(x.device).forward(x)

    # -- feature taps ----------------------------------------------------------
    def _tap_engine(self, device: torch.device, layers):
        """(engine, missing): the engine variant that stores every tapped slot.  The default program
        evaluates the upsampled half of the last decoder conv at low resolution and never builds that
        concat tensor; a tap there uses the single-launch variant (ANX_FLAG_NO_UPCONV)."""
        valid = []
        n_slots = len(self.module.model)
        for i in layers:            # the reference matches slots with `layer_id in layers`: any integer-like entry counts
            try:
                k = operator.index(i)
            except TypeError:
                continue
            if 0 <= k < n_slots:
                valid.append(k)
        eng, missing = None, valid
        for flags in (0, _lib.FLAG_NO_UPCONV):
            eng = self.engine_for(device, flags)
            table = eng.tap_table()
            missing = [i for i in valid if i not in table]
            if not missing:
                break
        return eng, missing

    def untappable(self, device: torch.device, layers):
        return self._tap_engine(device, layers)[1]

    def forward_taps(self, x: torch.Tensor, layers, encode_only: bool):
        eng, missing = self._tap_engine(x.device, layers)
        assert not missing, missing
        return eng.forward_taps(x, list(layers), encode_only)


def binding_for(module: nn.Module, cfg: dict) -> ModuleBinding:
    """The (lazily created) engine binding of a module; never stored on the module itself."""
    b = _BINDINGS.get(module)
    if b is None:
        b = ModuleBinding(module, cfg)
        _BINDINGS[module] = b
    return b


def invalidate(module: nn.Module) -> None:
    """Call after editing parameters through ``.data`` (or any other way that bypasses autograd's version
    counters) to make the next engine forward re-pack the weights immediately instead of at the next
    periodic content check."""
    b = _BINDINGS.get(module)
    if b is not None:
        b.invalidate()
