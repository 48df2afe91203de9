#!/usr/bin/env python
"""Headline benchmark: volumes/s of the anatomix 6M U-Net forward on 128^3 volumes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one forward of the engine over one batch of synthetic volumes
(BASELINE.json configs[1]: 6M U-Net, batch 8 x 1 x 128^3, bf16 operands / fp32
accumulate, fp32 in and out).  With N > 1 (launched by torchrun, one rank per
GPU) every rank runs its own batch (weak scaling, no data-path collective in the
timed value; the optional NCCL feature all-gather is timed separately and
reported under "allgather").  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's own CPU path for the same metric: the
reference is pure Python over torch ATen, so the arm runs the oracle's
torch-functional port (same ATen operators, all host threads), one 128^3 volume
per step.
"""
from __future__ import annotations

import argparse
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
VOL = 128
BATCH = 8
CONV_GFLOP_PER_VOL = 346.986          # sum 2*27*Cin*Cout*DHW, SURVEY.md appendix A
ALGO_MB_PER_VOL = 1025.0              # fused-minimum HBM bytes, SURVEY.md section 8(d)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, source="fallback")


def weights():
    """Real released 6M weights when the fixture is present, else seeded random."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "anatomix_6m_state.npz")
    if os.path.exists(path):
        z = np.load(path)
        return {k: torch.from_numpy(z[k]) for k in z.files}, "anatomix.pth (released 6M weights)"
    from oracle.unet_oracle import random_state
    return random_state(CFG_6M, 0), "seeded random weights"


def ncu_traffic_bytes():
    """DRAM bytes (read + write) of the conv kernels' launches in one 8x128^3 forward, from the committed
    ncu --set full capture (profiles/r1_traffic.json); None when that file is missing."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get("conv3_umma_kernel_dram_bytes_per_step")


def synth(n, seed):
    return torch.rand(n, 1, VOL, VOL, VOL, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from anatomix_b200.engine import Engine
    state, wsrc = weights()
    eng = Engine(CFG_6M, dev)
    eng.load_state(state)
    B = args.batch
    peaks = load_peaks()

    # inputs: rotate over enough distinct batches that their bytes exceed the 126 MB L2
    n_in = 3
    xs = [synth(B, 100 * rank + i).to(dev) for i in range(n_in)]
    out = torch.empty((B, 16, VOL, VOL, VOL), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        eng.forward(xs[i % n_in], out=out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        eng.forward(xs[i % n_in], out=out)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary()

    # end to end through the host-buffer C-ABI call: pinned input -> H2D -> forward -> D2H
    x_host = [synth(B, 7 + i).pin_memory() for i in range(2)]
    y_host = torch.empty((B, 16, VOL, VOL, VOL), dtype=torch.float32).pin_memory()
    dev_in = torch.empty((B, 1, VOL, VOL, VOL), dtype=torch.float32, device=dev)
    e2e_steps = max(3, min(args.steps, 10))
    eng.forward_host(x_host[0], y_host, dev_in, out)
    barrier()
    ev0.record()
    for i in range(e2e_steps):
        eng.forward_host(x_host[i % 2], y_host, dev_in, out)
    ev1.record()
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)

    # per-launch device times (CUDA events around every launch of one forward), averaged
    reps, acc = 5, {}
    order = []
    for r in range(reps):
        for name, t in eng.profile(xs[r % n_in]):
            if name not in acc:
                acc[name] = 0.0
                order.append(name)
            acc[name] += t / reps
    conv_ms = sum(t for n, t in acc.items() if "conv" in n and not n.startswith("conv0_"))   # conv3_umma_kernel launches
    fwd_ms = sum(acc.values())

    times = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e = times.tolist()

    allgather = None
    if world > 1:
        gathered = torch.empty((world * B, 16, VOL, VOL, VOL), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, out)
        barrier()
        ev0.record()
        for _ in range(3):
            dist.all_gather_into_tensor(gathered, out)
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allgather = {"ms": t.item(), "bytes_per_rank": out.numel() * 4,
                     "volumes_per_s_with_gather": world * B / ((ms / args.steps + t.item()) / 1e3)}
        # NCCL gather of step k on a side stream while step k+1 computes (double-buffered outputs)
        comm = torch.cuda.Stream(device=dev)
        outs = [out, torch.empty_like(out)]
        gath = [gathered, torch.empty_like(gathered)]
        done = [None, None]
        def pipelined(nsteps):
            for i in range(nsteps):
                b = i & 1
                if done[b] is not None:
                    torch.cuda.current_stream(dev).wait_event(done[b])   # gather that read outs[b] has finished
                eng.forward(xs[i % n_in], out=outs[b])
                ready = torch.cuda.Event()
                ready.record()
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    dist.all_gather_into_tensor(gath[b], outs[b])
                    done[b] = torch.cuda.Event()
                    done[b].record()
            torch.cuda.current_stream(dev).wait_stream(comm)
        pipelined(2)
        barrier()
        ev0.record()
        pipelined(args.steps)
        ev1.record()
        barrier()
        tp = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        allgather["overlapped_ms_per_step"] = tp.item() / args.steps
        allgather["volumes_per_s_overlapped_gather"] = world * B * args.steps / (tp.item() / 1e3)
        del outs, gath
        # the same gather fused into the last conv (epilogue stores into every peer's buffer over NVLink)
        try:
            from anatomix_b200.dist import FusedGatherExtractor
            fused = FusedGatherExtractor(eng)
            for i in range(2):
                fused.extract(xs[i % n_in])
            barrier()
            ev0.record()
            for i in range(args.steps):
                fused.extract(xs[i % n_in])
            ev1.record()
            barrier()
            tf = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            allgather["fused_ms_per_step"] = tf.item() / args.steps
            allgather["volumes_per_s_fused_gather"] = world * B * args.steps / (tf.item() / 1e3)
        except Exception as ex:          # symmetric memory unavailable on this box
            allgather["fused_error"] = str(ex)[:200]

    if rank == 0:
        vols = world * B * args.steps
        value = vols / (ms / 1e3)
        # stem conv (CUDA cores) excluded from the tensor-core roofline: 1.81 GFLOP of 346.99
        tc_flops = (CONV_GFLOP_PER_VOL - 1.812) * 1e9 * B
        achieved = tc_flops / (conv_ms / 1e3) / 1e12
        line = {
            "metric": "volumes/sec (128^3 1->16ch UNet forward)", "value": value, "unit": "volumes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"anatomix 6M UNet, batch {B}x128^3 bf16 on 1xB200 per rank (BASELINE configs[1])",
                       "weights": wsrc, "l2": f"{n_in} distinct input batches rotated; activations per step "
                       f"({eng.workspace_bytes(B, VOL, VOL, VOL) >> 20} MiB) exceed the 126 MB L2",
                       "parallelism": f"batch-sharded x{world}, no data-path collective in the timed value"},
            "e2e": {"value": world * B * e2e_steps / (ms_e2e / 1e3), "unit": "volumes/s",
                    "h2d_bytes_per_step": B * VOL ** 3 * 4, "d2h_bytes_per_step": B * 16 * VOL ** 3 * 4},
            "gpu_launches": eng.launches_per_forward(B, VOL, VOL, VOL) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv3_umma_kernel + conv3_rows_kernel, the two tcgen05 conv kernels "
                                                      "(%d launches per forward)" % sum(
                             1 for n in order if "conv" in n and not n.startswith("conv0_")),
                         "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tf_sustained"], "peak_source": peaks["source"] + " (sustained)",
                         "traffic": ncu_traffic_bytes(),
                         "whole_forward": {"ms_sum_of_launches": fwd_ms,
                                           "hbm_gbs_algorithmic": ALGO_MB_PER_VOL * 1e6 * B / (fwd_ms / 1e3) / 1e9,
                                           "hbm_peak": peaks["hbm"],
                                           "ceiling_vol_s": 4068.0, "frac_of_ceiling": value / world / 4068.0}},
            "launch_ms": {n: round(acc[n], 4) for n in order},
        }
        if allgather:
            line["allgather"] = allgather
        if not args.no_cpu_baseline and world == 1:
            torch.set_num_threads(os.cpu_count())
            ts = cpu_port_time(4, 1, state)
            line["cpu_baseline"] = {"value": len(ts) / sum(ts), "unit": "volumes/s",
                                    "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "4 single-volume 128^3 forwards (1 warm-up) of the oracle's "
                                              "torch-ATen port of the reference CPU path"}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
