"""Per-source-line (deepest frame inside the .cu file) instruction counts and stall samples from an ncu SASS export
joined with nvdisasm --print-line-info-inline.  usage: ncu_regions2.py sass.csv disasm.txt kernel_substr file.cu [top]"""
import csv, re, sys, collections
csvp, sassp, kern, cu = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
rows = list(csv.reader(open(csvp)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
cols = {n: H.index(n) for n in ["Address", "Source", "Instructions Executed", "# Samples", "stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_sleep", "stall_mio", "stall_branch_resolving"]}
inst = []
for r in rows[hdr + 1:]:
    if len(r) <= cols["# Samples"]: continue
    g = lambda n: int(r[cols[n]] or 0)
    inst.append((int(r[cols["Address"]], 16), r[cols["Source"]], g("Instructions Executed"), g("# Samples"), g("stall_long_sb"), g("stall_barrier"), g("stall_short_sb"), g("stall_wait"), g("stall_sleep"), g("stall_mio")))
base = inst[0][0]
cuname = cu.split("/")[-1]
off2line = {}
frames = []; infn = False; fresh = True
for ln in open(sassp):
    if ln.startswith(".text."):
        infn = kern in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh: frames = []; fresh = False
        frames.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        fresh = True
        deep = next((f for f in frames if f[0] == cuname), ("?", 0))
        off2line[int(m.group(1), 16)] = deep[1]
agg = collections.defaultdict(lambda: [0] * 8)
for a, txt, n, s, l, b, sh, w, sl, mio in inst:
    k = off2line.get(a - base, 0)
    v = agg[k]
    for i, x in enumerate((n, s, l, b, sh, w, sl, mio)): v[i] += x
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total warp-instr", tot, "samples", tots)
src = open(cu).read().splitlines()
if len(sys.argv) > 6:   # regions: "name:lo-hi,name:lo-hi"
    for spec in sys.argv[6].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        v = [sum(agg[k][i] for k in agg if lo <= k <= hi) for i in range(8)]
        print(f"{name:14s} instr {100*v[0]/tot:5.1f}%  samples {100*v[1]/tots:5.1f}%  long_sb {100*v[2]/tots:5.1f} barrier {100*v[3]/tots:5.1f} short_sb {100*v[4]/tots:5.1f} wait {100*v[5]/tots:5.1f} sleep {100*v[6]/tots:5.1f} mio {100*v[7]/tots:5.1f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = src[k - 1].strip()[:80] if 0 < k <= len(src) else ""
    print(f"L{k:5d} instr {100*v[0]/tot:5.1f}%  samp {100*v[1]/tots:5.1f}% (lsb {100*v[2]/tots:4.1f} bar {100*v[3]/tots:4.1f} ssb {100*v[4]/tots:4.1f} wait {100*v[5]/tots:4.1f})  {text}")
