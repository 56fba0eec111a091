"""Join an `ncu --page source --csv` SASS export with `nvdisasm --print-line-info` to get per-source-line instruction
counts and stall samples.  usage: ncu_lines.py <sass.csv> <nvdisasm.txt> <mangled kernel substring> <file.cu> [top]"""
import csv, re, sys, collections
csvp, sassp, kern, cu = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
rows = list(csv.reader(open(csvp)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ia, ii, istall = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
inst = []
for r in rows[hdr + 1:]:
    if len(r) <= istall: continue
    inst.append((int(r[ia], 16), int(r[ii] or 0), int(r[istall] or 0), r[1]))
base = inst[0][0]
# nvdisasm: find the function, map offset -> innermost line (in cu) using inline context lines
off2line = {}
cur = None; infn = False
for ln in open(sassp):
    if ln.startswith(".text."):
        infn = kern in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        # with inlining, nvdisasm prints "inlined at" chain; take the line of the cu file (outermost non-header)
        cur = (m.group(1), int(m.group(2)), m.group(3))
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
cnt = collections.Counter(); st = collections.Counter()
tot = 0; tots = 0
for a, n, s, txt in inst:
    k = off2line.get(a - base, ("?", 0, ""))
    key = (k[0].split("/")[-1], k[1])
    cnt[key] += n; st[key] += s; tot += n; tots += s
print("total warp-instr", tot, "samples", tots)
src = open(cu).read().splitlines()
hsrc = {}
for (f, l), n in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    text = src[l - 1].strip()[:90] if f == cu.split("/")[-1] and 0 < l <= len(src) else ""
    print(f"{n:11d} {100*n/tot:5.1f}%  samp {100*st[(f,l)]/max(tots,1):5.1f}%  {f}:{l}  {text}")
if len(sys.argv) > 6:
    want = set(int(x) for x in sys.argv[6].split(","))
    for a, n, s, txt in inst:
        k = off2line.get(a - base, ("?", 0, ""))
        if k[1] in want and n > 0:
            print(f"{a-base:6x} {k[1]:5d} n={n:9d} s={s:5d}  {txt.strip()[:100]}")
