#!/usr/bin/env python3
"""Per-device-function share of the warp-state samples of one kernel in an .ncu-rep (development aid).
usage: ncu_by_function.py report.ncu-rep library.so [kernel_mangled_prefix]
The noinline phases of the chain engine are separate symbols inside the kernel's text section; their (offset, size) come
from the cubin's symbol table, the samples from ncu's SASS page (addresses are consecutive from the section start)."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, so = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "_Z7k_chainN5fthmc9ChainArgsE"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sym = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
funcs = []
for line in sym.splitlines():
    p = line.split()
    if len(p) >= 8 and p[3] == "FUNC" and p[7].startswith("$" + kern + "$"):
        size = int(p[2], 16) if p[2].startswith("0x") else int(p[2])
        funcs.append((int(p[1], 16), size, p[7].split("$")[2]))
funcs.sort()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h, data = rows[1], rows[2:]
ia, isamp, iex, isrc = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
base = int(data[0][ia], 16)
def name_of(off):
    for o, s, n in funcs:
        if o <= off < o + s:
            n = re.sub(r"^_ZN5fthmc6EngineI\d+\w+?ExecE\d+", "", n)
            return n[:34]
    return "(kernel body / inlined)"
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
tot = 0
for r in data:
    n = name_of(int(r[ia], 16) - base)
    s = int(r[isamp] or 0)
    a = agg[n]; a[0] += s; a[1] += int(r[iex] or 0); tot += s
    for i in stall:
        a[2][h[i][6:]] += int(r[i] or 0)
    op = r[isrc].split()
    if op:
        a[3][op[1] if op[0].startswith("@") and len(op) > 1 else op[0]] += int(r[iex] or 0)
print(f"{'function':36s} samples%  warp-inst   top stalls | top opcodes by executed count")
for n, (s, ex, st, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if s < 0.003 * tot:
        continue
    print(f"{n:36s} {100 * s / tot:6.1f}%  {ex:9.3g}   " + " ".join(f"{k}:{100 * v / max(1, s):.0f}%" for k, v in st.most_common(4))
          + " | " + " ".join(f"{k}:{100 * v / max(1, ex):.0f}%" for k, v in ops.most_common(6)))
