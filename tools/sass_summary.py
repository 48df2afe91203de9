"""Per-kernel counts of the SASS mnemonics that prove Blackwell-native code paths (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld / st -> LDTM / STTM, TMA -> UTMALDG / UTMASTG / UBLKCP) in the shipped library.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "anatomix_b200", "lib", "libanatomix_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "SYNCS", "HMMA", "STG", "LDG", "ST.E", "LD.E"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# {os.path.relpath(LIB, ROOT)}  ({os.path.getsize(LIB)} bytes), cuobjdump -sass, sm_100a")
print(f"# {'kernel':100s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS[:7]))
tot = collections.Counter()
for fn in order:
    c = counts[fn]
    name = re.sub(r"\(.*", "", demangle(fn))[:100]
    print(f"  {name:100s} {c['instructions']:6d} " + " ".join(f"{c[k]:7d}" for k in KEYS[:7]))
    tot.update(c)
print(f"  {'TOTAL':100s} {tot['instructions']:6d} " + " ".join(f"{tot[k]:7d}" for k in KEYS[:7]))
print(f"# legacy tensor path (mma.sync / wmma -> HMMA): {tot['HMMA']} instructions")
