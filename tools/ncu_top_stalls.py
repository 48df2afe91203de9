"""Top stall sites of each kernel in an `ncu --page source --csv --print-source sass` dump.
Usage: python tools/ncu_top_stalls.py dump.csv [top_n] [kernel ordinal ...]"""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = set(int(a) for a in sys.argv[3:])
kernels, cur = [], None
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
for k, K in enumerate(kernels):
    if only and k not in only:
        continue
    h = K["hdr"]
    si, ii = h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si] or 0) for r in K["rows"])
    print(f"=== kernel {k}: {K['name'][:60]}  instructions {len(K['rows'])}  samples {tot}")
    agg = {}
    for r in K["rows"]:
        for i, c in stall_cols:
            agg[c] = agg.get(c, 0) + int(r[i] or 0)
    print("   stall totals:", {c: v for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    ranked = sorted(range(len(K["rows"])), key=lambda j: -int(K["rows"][j][si] or 0))[:top]
    for j in sorted(ranked):
        r = K["rows"][j]
        st = sorted(((int(r[i] or 0), c) for i, c in stall_cols), reverse=True)[:2]
        print(f"   {j:5d} {int(r[si]):6d} {100.0*int(r[si])/max(tot,1):5.1f}%  exec {r[ii]:>9}  {r[1].strip()[:70]:70s} {st}")
