#!/usr/bin/env python
"""Headline benchmark: volumes/s of the anatomix U-Net forward on 128^3 volumes (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one forward of the engine over one batch of synthetic volumes.  Rank 0 prints ONE JSON line.

N = 1 (BASELINE configs[1]): 6M U-Net, batch 8 x 1 x 128^3, bf16 operands / fp32 accumulate, fp32 in and out.
    `value` is device-resident, `e2e` goes through the host-buffer C-ABI call.  Sub-records under "configs":
    configs[2] (94M `anatomix-dev`, 4 x 128^3, with its own roofline), the production call shape (batch 2),
    one 1 x 512^3 volume on one GPU, a 256^3 sliding-window scan.
N > 1 (launched by torchrun, one rank per GPU; BASELINE configs[3]): every rank runs its own batch of 8 and the
    fp32 features of all ranks are all-gathered to all ranks INSIDE the timed region (`value`); the same loop
    without the gather is reported beside it, and "gather" lists every gather implementation.  "halo_512" is
    configs[4] (one 1 x 512^3 volume, depth-halo partition over the N GPUs) and "parity" re-checks the multi-GPU
    paths against the single-GPU engine on the box the numbers come from.

`--impl reference` times the reference's own CPU path for the same metric: the reference is pure Python over torch
ATen, so the arm runs the oracle's torch-functional port (same ATen operators, all host threads), one 128^3 volume
per step.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG_6M = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
CFG_94M = dict(dimension=3, input_nc=1, output_nc=32, num_downs=5, ngf=32, norm="instance", pooling="Avg",
               interp="trilinear", norm_eps=1e-2)
VOL = 128
BATCH = 8
# algorithmic work per 128^3 volume
GFLOP_6M, GFLOP_6M_STEM, MB_6M, CEIL_6M = 346.986, 1.812, 1025.0, 4068.0
GFLOP_94M, GFLOP_94M_STEM, MB_94M, CEIL_94M = 1418.748, 3.624, 2209.5, 1101.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def weights():
    """Real released 6M weights when the fixture is present, else seeded random."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "anatomix_6m_state.npz")
    if os.path.exists(path):
        z = np.load(path)
        return {k: torch.from_numpy(z[k]) for k in z.files}, "anatomix.pth (released 6M weights)"
    from oracle.unet_oracle import random_state
    return random_state(CFG_6M, 0), "seeded random weights"


def weights_94m():
    """The 94M checkpoint is Hub-only: the reference constructor's default init, seed 0 (as golden G4 / G8)."""
    from anatomix_b200 import Unet
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        return Unet(**CFG_94M).state_dict()


def ncu_traffic_bytes():
    """DRAM bytes (read + write) of the tcgen05 conv launches of one 8x128^3 forward, from the newest committed
    ncu --set full capture under profiles/; None when missing.  Third value: the per-launch list of that capture."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            d = json.load(open(path))
            return d.get("conv3_umma_kernel_dram_bytes_per_step"), name, d.get("per_launch")
    return None, None, None


# SURVEY Appendix A.1, MB per volume of the layers it marks memory-bound (bf16 intermediates, fp32 network input / output)
ALGO_MB_MEMORY_BOUND = {"conv0": 75.5, "conv3": 134.2, "conv6": 134.2, "conv62": 134.2, "conv65": 201.3}


def hbm_side(acc, order, per_launch, peaks, batch):
    """The memory-bound stages (the level-0 launches): DRAM bytes per launch from the committed ncu
    capture (a property of the kernel and the shape, batch 8) over the launch time measured live in this run."""
    tc = [n for n in order if "conv" in n]              # the tcgen05 launches of one forward, in launch order
    if not per_launch or len(per_launch) != len(tc) or batch != BATCH:
        return None
    out = {}
    for n, pl in zip(tc, per_launch):
        if n.endswith("_L0") and acc[n] > 0:
            gbs = pl["dram_gb"] / (acc[n] / 1e3)
            out[n] = {"dram_gb_ncu": pl["dram_gb"], "ms_live": round(acc[n], 4), "gbs": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3)}
            mb = ALGO_MB_MEMORY_BOUND.get(n.split("_")[0])
            if mb:      # the same launch against SURVEY's fused-minimum bytes of that layer
                out[n]["algorithmic_gb"] = round(mb * batch / 1e3, 4)
                out[n]["frac_of_hbm_peak_algorithmic"] = round(mb * 1e6 * batch / (acc[n] / 1e3) / 1e9 / peaks["hbm"], 3)
    return {"hbm_peak_gbs": peaks["hbm"], "launches": out}


def synth(n, seed, size=VOL):
    return torch.rand(n, 1, size, size, size, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled WHILE the timed region runs: NVML every few
    milliseconds (falls back to polling nvidia-smi when pynvml is unavailable)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.bits, self.max_mhz, self.stop_flag = index, [], 0, None, threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES remapping via the UUID of the torch device
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            handle = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    handle = h
            self.handle = handle or pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml:
                    self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
                    self.bits |= int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    self.stop_flag.wait(0.004)
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index),
                                          "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.sm.append(float(out[0]))
                    self.max_mhz = float(out[1])
                    self.stop_flag.wait(0.2)
            except Exception:
                self.stop_flag.wait(0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        reasons = [name for bit, name in self.REASONS.items() if self.bits & bit]
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def cpu_port_time(steps, warmup, state):
    """Seconds per 128^3 volume of the reference's CPU path (oracle port)."""
    from oracle.unet_oracle import unet_forward
    x = synth(1, 0)
    for _ in range(warmup):
        unet_forward(CFG_6M, state, x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        unet_forward(CFG_6M, state, x)
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    state, wsrc = weights()
    ts = cpu_port_time(args.steps, max(1, min(args.warmup, 2)), state)
    total = sum(ts)
    v = len(ts) / total
    line = {
        "impl": "reference", "metric": "volumes/sec (128^3 1->16ch UNet forward)", "value": v, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "anatomix 6M UNet forward, 1x(1,128,128,128) fp32 per step on the host CPU "
                               "(reference PyTorch ATen path, oracle port)", "weights": wsrc},
        "cpu_baseline": {"value": v, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{len(ts)} single-volume forwards of the same network and input size"},
        "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Timer:
    """CUDA-event timing on the current stream of `dev`, max over ranks."""

    def __init__(self, dev, world, dist):
        self.dev, self.world, self.dist = dev, world, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def run(self, fn, steps, warmup):
        """ms per call of fn(i) over `steps` calls after `warmup` untimed ones (barrier + synchronize on both sides)."""
        for i in range(warmup):
            fn(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps


def launch_profile(eng, xs, reps=5):
    acc, order = {}, []
    for r in range(reps):
        for name, t in eng.profile(xs[r % len(xs)]):
            if name not in acc:
                acc[name] = 0.0
                order.append(name)
            acc[name] += t / reps
    return acc, order


def is_tc_conv(name):       # launches of the two tcgen05 conv kernels (the stem has its own kernel)
    return "conv" in name and not name.startswith("conv0_")


def roofline_block(acc, order, gflop_vol, gflop_stem, mb_vol, ceiling, batch, vol_s, peaks, traffic=None, traffic_src=None,
                   per_launch=None):
    conv_ms = sum(t for n, t in acc.items() if is_tc_conv(n))
    fwd_ms = sum(acc.values())
    achieved = (gflop_vol - gflop_stem) * 1e9 * batch / (conv_ms / 1e3) / 1e12
    return {"bound": "tensor",
            "kernel": "conv3_umma_kernel + conv3_rows_kernel, the tcgen05 conv kernels (%d launches per forward)"
                      % sum(1 for n in order if is_tc_conv(n)),
            "achieved": achieved, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": achieved / peaks["tf"],
            "peak_source": peaks["source"] + " bf16_tflops (burst: the timed region is short and runs at full SM clock)",
            "frac_of_sustained_peak": achieved / peaks["tf_sustained"], "sustained_peak": peaks["tf_sustained"],
            "traffic": traffic, "traffic_source": traffic_src,
            "memory_bound_launches": hbm_side(acc, order, per_launch, peaks, batch),
            "whole_forward": {"ms_sum_of_launches": fwd_ms,
                              "hbm_gbs_algorithmic": mb_vol * 1e6 * batch / (fwd_ms / 1e3) / 1e9, "hbm_peak": peaks["hbm"],
                              "ceiling_vol_s": ceiling, "frac_of_ceiling": vol_s / ceiling}}


# ----------------------------------------------------------------------------------------------- N = 1 sub-records
def single_gpu_configs(dev, tm, state, peaks, steps):
    from anatomix_b200.engine import Engine
    out = {}
    # the production call shape: MONAI's sliding-window predictor hands the model [2, 1, 128, 128, 128] windows
    # (reference convex_adam_utils.py:202-210, sw_batch_size = 2)
    eng = Engine(CFG_6M, dev)
    eng.load_state(state)
    xs = [synth(2, 50 + i).to(dev) for i in range(3)]
    o2 = torch.empty((2, 16, VOL, VOL, VOL), dtype=torch.float32, device=dev)
    ms = tm.run(lambda i: eng.forward(xs[i % 3], out=o2), max(steps, 20), 3)
    out["6m_2x128_production_shape"] = {"ms_per_step": ms, "volumes_per_s": 2e3 / ms,
                                        "note": "batch 2 = sw_batch_size of the registration caller"}
    del xs, o2
    # one 1 x 512^3 volume on ONE GPU: the single-device reference point of configs[4]
    try:
        x5 = torch.rand(1, 1, 512, 512, 512, device=dev)
        o5 = torch.empty((1, 16, 512, 512, 512), dtype=torch.float32, device=dev)
        ms = tm.run(lambda i: eng.forward(x5, out=o5), 3, 1)
        out["6m_1x512_one_gpu"] = {"ms_per_volume": ms, "equiv_128_volumes_per_s": 64e3 / ms,
                                   "workspace_gib": eng.workspace_bytes(1, 512, 512, 512) / 2 ** 30}
        del x5, o5
    except Exception as ex:
        out["6m_1x512_one_gpu"] = {"error": str(ex)[:200]}
    eng._workspaces.clear()
    torch.cuda.empty_cache()
    # a 256^3 scan through the sliding-window inferer in the registration setting (128^3 windows, overlap 0.8,
    # gaussian blend: reference convex_adam_utils.py:202-219), windows fed to the engine in batches of 8
    try:
        from anatomix_b200.sliding import sliding_window_features
        scan = torch.rand(1, 1, 256, 256, 256, device=dev)
        ms = tm.run(lambda i: sliding_window_features(scan, (128, 128, 128), 8, eng.forward, overlap=0.8,
                                                      mode="gaussian", sigma_scale=0.25), 2, 1)
        out["6m_sliding_scan_256"] = {"ms_per_scan": ms, "windows": 343, "windows_per_s": 343e3 / ms}
        del scan
    except Exception as ex:
        out["6m_sliding_scan_256"] = {"error": str(ex)[:200]}
    del eng
    torch.cuda.empty_cache()
    # BASELINE configs[2]: anatomix-dev 94M (InstanceNorm, AvgPool, trilinear), batch 4 x 128^3
    try:
        e94 = Engine(CFG_94M, dev)
        e94.load_state(weights_94m())
        xs = [synth(4, 60 + i).to(dev) for i in range(3)]
        o94 = torch.empty((4, 32, VOL, VOL, VOL), dtype=torch.float32, device=dev)
        ms = tm.run(lambda i: e94.forward(xs[i % 3], out=o94), max(steps // 2, 10), 3)
        acc, order = launch_profile(e94, xs, 3)
        vs = 4e3 / ms
        out["94m_4x128"] = {
            "workload": "anatomix-dev 94M UNet, batch 4x128^3 on 1xB200 (BASELINE configs[2]); fp16 operands / fp32 "
                        "accumulate; seeded default-init weights (the released checkpoint is Hub-only)",
            "ms_per_step": ms, "volumes_per_s": vs, "gpu_launches_per_step": e94.launches_per_forward(4, VOL, VOL, VOL),
            "roofline": roofline_block(acc, order, GFLOP_94M, GFLOP_94M_STEM, MB_94M, CEIL_94M, 4, vs, peaks),
            "launch_ms": {n: round(acc[n], 4) for n in order},
            "non_conv_ms": sum(t for n, t in acc.items() if not is_tc_conv(n))}
        del e94, xs, o94
    except Exception as ex:
        out["94m_4x128"] = {"error": str(ex)[:300]}
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------- N > 1 sub-records
def multi_gpu_parity(dev, rank, world, dist, eng, state):
    """Multi-GPU paths against the single-GPU engine, on the box the numbers come from (small shapes)."""
    from anatomix_b200.dist import FeatureGather, slab_bounds
    from anatomix_b200.halo import DepthSlabExtractor
    from anatomix_b200.engine import Engine
    res = {}

    def agree(flag):                       # True only if every rank saw True
        t = torch.tensor([1 if flag else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def worst(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    n = 2
    full = torch.rand(world * n, 1, 32, 32, 128, generator=torch.Generator().manual_seed(77)).to(dev)
    mine = full[rank * n:(rank + 1) * n].contiguous()
    ref = eng.forward(full)
    try:
        gathered = torch.empty_like(ref)
        dist.all_gather_into_tensor(gathered, eng.forward(mine))
        res["nccl_gather_equals_local_forward"] = agree(torch.equal(gathered, ref))
        push = FeatureGather(eng, mode="push", payload="f32").extract(mine)
        torch.cuda.synchronize(dev)
        res["push_gather_equals_local_forward"] = agree(torch.equal(push, ref))
        fused = FeatureGather(eng, mode="fused", payload="f32").extract(mine)
        torch.cuda.synchronize(dev)
        res["fused_gather_equals_nccl_gather"] = agree(torch.equal(fused, gathered))
        ref16 = eng.forward_cl16(full)
        ok = True
        for mode in ("push", "fused"):
            g16 = FeatureGather(eng, mode=mode, payload="cl16").extract(mine)
            torch.cuda.synchronize(dev)
            ok = ok and torch.equal(g16, ref16)
        res["cl16_gathers_equal_local_16bit_forward"] = agree(ok)
        res["cl16_widened_equals_rounded_fp32"] = agree(torch.equal(eng.widen(ref16), ref.to(eng.storage_dtype).float()))
    except Exception as ex:
        res["gather_error"] = str(ex)[:300]
    try:
        depth = 32 * world
        vol = torch.rand(1, 1, depth, 64, 128, generator=torch.Generator().manual_seed(78))
        lo, hi = slab_bounds(depth, world, 4)[rank]
        want = eng.forward(vol.to(dev))[:, :, lo:hi]
        for in_engine, key in ((True, "halo_in_engine_max_abs_vs_single_gpu"), (False, "halo_nccl_steps_max_abs_vs_single_gpu")):
            slab = DepthSlabExtractor(CFG_6M, state, dev, in_engine_exchange=in_engine)
            got = slab.extract(vol)
            got = slab.extract(vol)                    # twice: sequence numbers / shells carry over between forwards
            torch.cuda.synchronize(dev)
            res[key] = worst((got - want).abs().max().item())
            del slab
        res["halo_slabs"] = world
        res["halo_volume"] = [depth, 64, 128]
    except Exception as ex:
        res["halo_error"] = str(ex)[:300]
    try:
        from oracle import unet_oracle as O
        cfg_in = dict(dimension=3, input_nc=1, output_nc=16, num_downs=2, ngf=16, norm="instance", pooling="Avg",
                      interp="trilinear", norm_eps=1e-2)
        st_in = O.random_state(cfg_in, seed=9)
        depth = 8 * world
        vol = torch.rand(1, 1, depth, 16, 24, generator=torch.Generator().manual_seed(79))
        lo, hi = slab_bounds(depth, world, 2)[rank]
        single = Engine(cfg_in, dev)
        single.load_state(st_in)
        want = single.forward(vol.to(dev))[:, :, lo:hi]
        got = DepthSlabExtractor(cfg_in, st_in, dev).extract(vol)
        torch.cuda.synchronize(dev)
        res["instance_norm_slab_rel_l2_vs_single_gpu"] = worst(((got - want).norm() / want.norm()).item())
    except Exception as ex:
        res["instance_norm_slab_error"] = str(ex)[:300]
    return res


def multi_gpu_gather(dev, rank, world, dist, tm, eng, xs, out, steps, ms_compute):
    """BASELINE configs[3]: every gather implementation with the gather INSIDE the timed region."""
    from anatomix_b200.dist import FeatureGather
    B = xs[0].shape[0]
    res = {"bytes_received_per_rank_per_step_fp32": (world - 1) * out.numel() * 4}

    def rate(ms):
        return world * B * 1e3 / ms
    # NCCL, gather after every forward on the same stream
    gathered = torch.empty((world * B,) + tuple(out.shape[1:]), dtype=torch.float32, device=dev)

    def nccl_seq(i):
        eng.forward(xs[i % len(xs)], out=out)
        dist.all_gather_into_tensor(gathered, out)
    ms = tm.run(nccl_seq, steps, 2)
    res["nccl_sequential"] = {"ms_per_step": ms, "volumes_per_s": rate(ms)}
    # NCCL on a side stream while the next step computes (double-buffered)
    comm = torch.cuda.Stream(device=dev)
    outs, gath, done = [out, torch.empty_like(out)], [gathered, torch.empty_like(gathered)], [None, None]

    def nccl_overlapped(i):
        b = i & 1
        if done[b] is not None:
            torch.cuda.current_stream(dev).wait_event(done[b])
        eng.forward(xs[i % len(xs)], out=outs[b])
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            dist.all_gather_into_tensor(gath[b], outs[b])
            done[b] = torch.cuda.Event()
            done[b].record()

    def drain():
        torch.cuda.current_stream(dev).wait_stream(comm)
    for i in range(2):
        nccl_overlapped(i)
    drain()
    tm.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        nccl_overlapped(i)
    drain()
    e1.record()
    tm.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["nccl_overlapped"] = {"ms_per_step": t.item(), "volumes_per_s": rate(t.item())}
    del outs, gath, gathered
    torch.cuda.empty_cache()

    # the engine's own gathers over peer-mapped buffers (no library collective)
    def pipelined(fg, widen_into=None):
        pending = []

        def step(i):
            pending.append(fg.submit(xs[i % len(xs)]))
            if len(pending) > 1:                    # consume step i - 1 while step i is in flight
                g = fg.wait(pending.pop(0))
                if widen_into is not None:
                    eng.widen(g, out=widen_into)

        def flush():
            while pending:
                g = fg.wait(pending.pop(0))
                if widen_into is not None:
                    eng.widen(g, out=widen_into)
        for i in range(2):
            step(i)
        flush()
        tm.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            step(i)
        flush()
        b.record()
        tm.barrier()
        tt = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt.item()
    for key, mode, payload, widen in (("push_fp32", "push", "f32", False), ("push_cl16", "push", "cl16", False),
                                      ("push_cl16_widened_to_fp32", "push", "cl16", True),
                                      ("fused_fp32", "fused", "f32", False), ("fused_cl16", "fused", "cl16", False)):
        try:
            fg = FeatureGather(eng, mode=mode, payload=payload, depth=2 if mode == "push" else 1)
            wide = torch.empty((world * B, 16, VOL, VOL, VOL), dtype=torch.float32, device=dev) if widen else None
            ms = pipelined(fg, wide)
            recv = (world - 1) * out.numel() * (4 if payload == "f32" else 2)
            res[key] = {"ms_per_step": ms, "volumes_per_s": rate(ms), "nvlink_ingress_gbs_per_rank": recv / ms / 1e6}
            del fg, wide
        except Exception as ex:
            res[key] = {"error": str(ex)[:300]}
        torch.cuda.empty_cache()
    res["compute_only_ms_per_step"] = ms_compute
    return res


def multi_gpu_halo(dev, rank, world, dist, tm, state, size=512):
    """BASELINE configs[4]: one 1 x size^3 volume, depth-halo partition over the ranks (slab inputs synthesised on
    the device: the upload of a slab is the same for every approach)."""
    from anatomix_b200.halo import DepthSlabExtractor
    from anatomix_b200.engine import Engine
    if size % (16 * world) or size // world < 32:
        return {"skipped": f"{size} planes do not split into {world} slabs of >= 32 planes aligned to 16"}
    d = size // world
    res = {"volume": [size, size, size], "slabs": world, "planes_per_slab": d}
    slab = DepthSlabExtractor(CFG_6M, state, dev)
    slab.engine.set_slab(rank > 0, rank < world - 1, size)
    x = torch.rand(1, 1, d + 2, size, size, device=dev)
    out = torch.empty((1, 16, d, size, size), dtype=torch.float32, device=dev)
    ms = tm.run(lambda i: slab.forward_slab(x, out), 5, 2)
    res["in_engine_exchange"] = {"ms_per_volume": ms, "equiv_128_volumes_per_s": (size / 128) ** 3 * 1e3 / ms,
                                 "exchanges_per_forward": sum(1 for s in slab.steps if s[1] >= 0)}
    res["workspace_gib_per_rank"] = slab.engine.workspace_bytes(1, d, size, size) / 2 ** 30
    # the same slab without any exchange: what the partition would cost with free communication
    plain = Engine(CFG_6M, dev)
    plain.load_state(state)
    xi = x[:, :, 1:-1].contiguous()
    ms0 = tm.run(lambda i: plain.forward(xi, out=out), 5, 2)
    res["slab_compute_only_ms"] = ms0
    res["exchange_overhead_ms"] = ms - ms0
    del plain, slab
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline measurement only (no sub-records)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    # Library chatter (NCCL prints its version banner on fd 1) must not share stdout with the ONE JSON
    # line: everything written to fd 1 from here on goes to stderr; the JSON goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    state, wsrc = weights()

    # CPU baseline FIRST (rank 0, N = 1 only), before any GPU work disturbs the host: >= 10 single-volume forwards
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count())
        ts = cpu_port_time(10, 2, state)
        cpu_baseline = {"value": len(ts) / sum(ts), "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": "10 single-volume 128^3 forwards (2 warm-ups) of the oracle's torch-ATen port of the "
                                  "reference CPU path, taken before the GPU legs",
                        "best": 1.0 / min(ts), "median": 1.0 / statistics.median(ts)}

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    tm = Timer(dev, world, dist)

    from anatomix_b200.engine import Engine
    eng = Engine(CFG_6M, dev)
    eng.load_state(state)
    B = args.batch
    peaks = load_peaks()

    # inputs: rotate over enough distinct batches that their bytes exceed the 126 MB L2
    n_in = 3
    xs = [synth(B, 100 * rank + i).to(dev) for i in range(n_in)]
    out = torch.empty((B, 16, VOL, VOL, VOL), dtype=torch.float32, device=dev)

    # ---- device-resident forward (the whole headline at N = 1; "without gather" at N > 1)
    for i in range(args.warmup):
        eng.forward(xs[i % n_in], out=out)
    tm.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        eng.forward(xs[i % n_in], out=out)
    ev1.record()
    tm.barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks_compute = sampler.summary()

    # ---- N > 1 headline: the fp32 feature all-gather inside the timed region (BASELINE configs[3])
    headline_gather = None
    if world > 1:
        try:
            from anatomix_b200.dist import FeatureGather
            fg = FeatureGather(eng, mode="push", payload="f32", depth=2)
            pending = []

            def step(i):
                pending.append(fg.submit(xs[i % n_in]))
                if len(pending) > 1:
                    fg.wait(pending.pop(0))
            for i in range(args.warmup):
                step(i)
            while pending:
                fg.wait(pending.pop(0))
            tm.barrier()
            sampler = ClockSampler(local)
            sampler.start()
            ev0.record()
            for i in range(args.steps):
                step(i)
            while pending:
                fg.wait(pending.pop(0))
            ev1.record()
            tm.barrier()
            headline_gather = {"ms_total": ev0.elapsed_time(ev1), "clocks": sampler.summary(),
                               "how": "anx_push_to_peers: copy-engine push of every rank's fp32 slice into all ranks' "
                                      "peer-mapped gather buffers, pipelined one step behind the convs"}
            del fg
            torch.cuda.empty_cache()
        except Exception as ex:                      # symmetric memory unavailable: NCCL on the same stream
            gathered = torch.empty((world * B, 16, VOL, VOL, VOL), dtype=torch.float32, device=dev)

            def nccl_step(i):
                eng.forward(xs[i % n_in], out=out)
                dist.all_gather_into_tensor(gathered, out)
            ms = tm.run(nccl_step, args.steps, args.warmup)
            headline_gather = {"ms_total": ms * args.steps, "clocks": clocks_compute,
                               "how": "ncclAllGather after every forward (peer-mapped buffers unavailable: %s)" % str(ex)[:120]}
            del gathered

    # ---- end to end through the host-buffer C-ABI call: pinned input -> H2D -> forward -> D2H
    x_host = [synth(B, 7 + i).pin_memory() for i in range(2)]
    y_host = torch.empty((B, 16, VOL, VOL, VOL), dtype=torch.float32).pin_memory()
    dev_in = torch.empty((B, 1, VOL, VOL, VOL), dtype=torch.float32, device=dev)
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e_single = tm.run(lambda i: eng.forward_host(x_host[i % 2], y_host, dev_in, out), e2e_steps, 1)
    # back-to-back calls the way a throughput caller issues them: two {device output, host output} sets alternate,
    # the download of call k overlaps the upload and convs of call k + 1 (anx_engine_forward_host_pipelined); the
    # timed region ends after anx_engine_host_wait + synchronize, i.e. with every result in host memory
    y_host2 = torch.empty_like(y_host).pin_memory()
    out2 = torch.empty_like(out)
    sets = [(y_host, out), (y_host2, out2)]

    def e2e_call(i):
        yh, od = sets[i % 2]
        eng.forward_host_pipelined(x_host[i % 2], yh, dev_in, od)
    e2e_call(0)
    eng.host_wait()
    tm.barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for i in range(e2e_steps):
        e2e_call(i)
    eng.host_wait()
    eb.record()
    tm.barrier()
    ms_e2e = ea.elapsed_time(eb)
    del y_host2, out2
    # the link itself: a plain pinned device -> host copy of one step's fp32 output (what bounds the e2e number)
    def d2h(i):
        y_host.copy_(out, non_blocking=True)
    ms_d2h = tm.run(d2h, 3, 1)
    pcie = {"d2h_gbs_plain_copy": out.numel() * 4 / ms_d2h / 1e6,
            "e2e_bound_volumes_per_s": world * B * 1e3 / ms_d2h,
            "note": "cudaMemcpyAsync of one step's fp32 features (pinned host memory), all ranks at once"}
    # the same call with the opt-in 16-bit channels-last payload (half the download)
    e2e_cl16 = None
    try:
        y16_host = torch.empty((B, VOL, VOL, VOL, 16), dtype=eng.storage_dtype).pin_memory()
        dev16 = torch.empty((B, VOL, VOL, VOL, 16), dtype=eng.storage_dtype, device=dev)
        ms16 = tm.run(lambda i: eng.forward_host_cl16(x_host[i % 2], y16_host, dev_in, dev16), e2e_steps, 1)
        e2e_cl16 = {"value": world * B * 1e3 / ms16, "unit": "volumes/s", "d2h_bytes_per_step": B * 16 * VOL ** 3 * 2,
                    "note": "opt-in 16-bit channels-last features (anx_engine_forward_host_ex, ANX_PAYLOAD_CL16); "
                            "NOT the drop-in fp32 output"}
        del y16_host, dev16
    except Exception as ex:
        e2e_cl16 = {"error": str(ex)[:200]}
    del x_host, y_host, dev_in

    # ---- per-launch device times (CUDA events around every launch of one forward), averaged
    acc, order = launch_profile(eng, xs)

    times = torch.tensor([ms_total, ms_e2e, headline_gather["ms_total"] if headline_gather else 0.0, ms_e2e_single],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_gather_total, ms_e2e_single = times.tolist()

    extras = {}
    if not args.no_extras and B == BATCH:
        if world == 1:
            extras["configs"] = single_gpu_configs(dev, tm, state, peaks, args.steps)
        else:
            extras["parity"] = multi_gpu_parity(dev, rank, world, dist, eng, state)
            extras["gather"] = multi_gpu_gather(dev, rank, world, dist, tm, eng, xs, out, max(5, args.steps // 2),
                                                ms_total / args.steps)
            try:
                extras["halo_512"] = multi_gpu_halo(dev, rank, world, dist, tm, state)
            except Exception as ex:
                extras["halo_512"] = {"error": str(ex)[:300]}

    if rank == 0:
        vols = world * B * args.steps
        value_compute = vols / (ms_total / 1e3)
        if world > 1:
            value, ms_step = vols / (ms_gather_total / 1e3), ms_gather_total / args.steps
            clocks = headline_gather["clocks"]
            par = (f"batch-sharded x{world} (8 volumes per rank per step) + all-gather of the fp32 [{world * B},16,128^3] "
                   f"features to every rank INSIDE the timed value: " + headline_gather["how"])
        else:
            value, ms_step, clocks = value_compute, ms_total / args.steps, clocks_compute
            par = "single GPU"
        traffic, traffic_src, per_launch = ncu_traffic_bytes()
        line = {
            "metric": "volumes/sec (128^3 1->16ch UNet forward)", "value": value, "unit": "volumes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"anatomix 6M UNet, batch {B}x128^3 bf16 on 1xB200 per rank "
                                   + ("(BASELINE configs[1])" if world == 1 else
                                      f"= batch {world * B}x128^3 batch-sharded across {world}xB200 with the feature "
                                      "all-gather (BASELINE configs[3])"),
                       "weights": wsrc, "l2": f"{n_in} distinct input batches rotated; activations per step "
                       f"({eng.workspace_bytes(B, VOL, VOL, VOL) >> 20} MiB) exceed the 126 MB L2",
                       "parallelism": par},
            "e2e": {"value": world * B * e2e_steps / (ms_e2e / 1e3), "unit": "volumes/s",
                    "h2d_bytes_per_step": B * VOL ** 3 * 4, "d2h_bytes_per_step": B * 16 * VOL ** 3 * 4,
                    "how": "anx_engine_forward_host_pipelined: pinned host input -> H2D -> forward -> D2H into pinned host "
                           "memory every step, back-to-back calls on two alternating output buffer sets; timed to "
                           "anx_engine_host_wait + synchronize (every result in host memory)"},
            "e2e_single_call": {"value": world * B * 1e3 / ms_e2e_single, "unit": "volumes/s",
                                "how": "anx_engine_forward_host: each call completes its download before the next starts"},
            "e2e_cl16": e2e_cl16,
            "e2e_link": pcie,
            "gpu_launches": eng.launches_per_forward(B, VOL, VOL, VOL) * args.steps,
            "clocks": clocks,
            "roofline": roofline_block(acc, order, GFLOP_6M, GFLOP_6M_STEM, MB_6M, CEIL_6M, B, value_compute / world,
                                       peaks, traffic, traffic_src, per_launch),
            "launch_ms": {n: round(acc[n], 4) for n in order},
        }
        if world > 1:
            line["value_without_gather"] = value_compute
            line["ms_per_step_without_gather"] = ms_total / args.steps
        line.update(extras)
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
