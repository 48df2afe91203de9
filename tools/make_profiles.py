"""Turns the raw outputs of tools/round_profiles.sh (gpurun_out/final/) into the committed profiles/r2_* files.
    python tools/make_profiles.py"""
import re
import collections, csv, json, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F, P = os.path.join(ROOT, "gpurun_out", "final") + "/", os.path.join(ROOT, "profiles") + "/"
for src, dst in (("bench_n1.json", "r2_bench_b200_n1.json"), ("bench_reference.json", "r2_bench_reference_cpu.json"),
                 ("rows.json", "r2_rows.json"), ("torch_gpu.json", "r2_torch_gpu_same_box.json"), ("tests.log", "r2_parity_values.txt")):
    if os.path.exists(F + src):
        shutil.copy(F + src, P + dst)

# launch list (every launch of `bench.py --steps 2 --warmup 3 --no-extras` with its device time)
rows = list(csv.reader(open(F + "launches.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]
ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
out = []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    out.append((r[0], r[ki], r[bi], r[gi], v))
with open(P + "r2_launches.csv", "w") as f:
    f.write("id,kernel,block,grid,duration_us\n")
    for o in out:
        f.write(f'{o[0]},"{o[1][:90]}","{o[2]}","{o[3]}",{o[4]:.3f}\n')
mine = [o for o in out if "anx::" in o[1] or "conv3_" in o[1] or "upsample2" in o[1] or "pool2" in o[1]]
bench = json.loads(open(F + "bench_n1.json").read().strip().splitlines()[-1])
per_fwd = bench["gpu_launches"] // bench["steps"]      # launches of one 8x128^3 forward of the 6M network
last = mine[4 * per_fwd:5 * per_fwd]         # the last forward of the timed headline region (3 warm-ups + 2 timed steps);
                                             # later launches belong to the e2e leg, which runs the batch in chunks of two
tot = sum(o[4] for o in last)
fam, cnt = collections.Counter(), collections.Counter()
for o in last:
    name = o[1].replace("void ", "").replace("anx::", "").split("(")[0]
    fam[name] += o[4]
    cnt[name] += 1
json.dump({"source": "ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: shares, not absolutes)",
           "forward": "last step of the timed device-resident region (batch 8)", "last_forward_us": tot, "launches": len(last),
           "by_kernel": {k: {"launches": cnt[k], "us": round(v, 1), "share": round(v / tot, 4)} for k, v in fam.most_common()}},
          open(P + "r2_launch_shares.json", "w"), indent=1)

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def table(raw, dst):
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:70] for i in idx])
    return rows, hdr


rows, hdr = table(F + "fwd_raw.csv", P + "r2_ncu_full_forward_umma_kernels.csv")
ki = hdr.index("Kernel Name")
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def traffic(r):
    return sum(float(r[hdr.index(n)]) * unit[rows[1][hdr.index(n)]] for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"))


data = rows[2:]
stem = [r for r in data if re.search(r"conv3_rows_kernel<\d+, 1[,>]", r[ki]) or "stem_umma" in r[ki]]   # <MODE, STEM = 1, ...>
conv = [r for r in data if r not in stem]
json.dump({"source": "ncu --set full --clock-control none over the tcgen05 launches of one 8x128^3 forward (tools/round_profiles.sh)",
           "conv_launches": len(conv), "conv3_umma_kernel_dram_bytes_per_step": sum(traffic(r) for r in conv),
           "stem_dram_bytes_per_step": sum(traffic(r) for r in stem), "algorithmic_bytes_per_step_whole_forward": 1025.0e6 * 8,
           "per_launch": [{"kernel": r[ki].replace("void ", "").split("(")[0][:40], "us": float(r[hdr.index("gpu__time_duration.sum")]),
                           "dram_gb": round(traffic(r) / 1e9, 4)} for r in data]},
          open(P + "r2_traffic.json", "w"), indent=1)
if os.path.exists(F + "fwd94_raw.csv"):
    table(F + "fwd94_raw.csv", P + "r2_ncu_94m_forward_kernels.csv")
print("launches in the last forward:", len(last), "conv launches captured:", len(conv), "stem:", len(stem))
