import csv, re, sys, collections
csvp, sassp, kern = sys.argv[1:4]
SRC = open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "pharmacoforge_b200", "csrc", "pf_tc_conv.cu")).read().splitlines()
def _ln(marker, start=0):
    for i in range(start, len(SRC)):
        if marker in SRC[i]: return i + 1
    raise KeyError(marker)
_m = [("producer", "__device__ void producer_role("), ("mma_role", "__device__ void mma_role("), ("mean_fn", "void lds128_if("),
      ("epi_setup", "__device__ void epilogue_role("), ("meta+xdiff", "slot_barrier(T);  // everyone is done with the previous tile"),
      ("gather_h", "// ---- gather h[src]"), ("gather_v+stage", "float Vu[24];"), ("EPI-A", "// ================= EPI-A"),
      ("EPI-B", "// ================= EPI-B"), ("mean_h", "if (g == 2) {  // segmented mean of the scalar"),
      ("EPI-C", "// ================= EPI-C"), ("mean_v", "float* ab = reinterpret_cast<float*>(stage);  // [128][kMeanPitchV]"),
      ("kernel", "edge_conv_tc_kernel(const Params p)"), ("end", "// K4 on the tensor cores")]
_l = [(n, _ln(k)) for n, k in _m]
regions = [(_l[i][1], _l[i + 1][1] - 1, _l[i][0]) for i in range(len(_l) - 1)]
def reg(l):
    for a,b,n in regions:
        if a<=l<=b: return n
    return None
rows = list(csv.reader(open(csvp)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ii, istall = H.index("Instructions Executed"), H.index("# Samples")
inst = [(int(r[0],16), int(r[ii] or 0), int(r[istall] or 0), r[1]) for r in rows[hdr+1:] if len(r)>istall]
base = inst[0][0]
off2 = {}; cur=None; infn=False; lastreg="kernel"
for ln in open(sassp):
    if ln.startswith(".text."): infn = kern in ln; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if m.group(1).endswith("pf_tc_conv.cu"):
            r = reg(int(m.group(2)))
            if r: lastreg = r
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: off2[int(m.group(1),16)] = (lastreg, cur)
cnt=collections.Counter(); st=collections.Counter(); spin=collections.Counter(); ops=collections.defaultdict(collections.Counter)
for a,n,s,txt in inst:
    r,cur = off2.get(a-base,("?",None))
    isspin = cur and ((cur[0]=="pf_tc.cuh" and cur[1] in (43,38,39,40)) or "YIELD" in txt or "TRYWAIT" in txt)
    if isspin: spin[r]+=n; st[r+"/wait"]+=s
    else: cnt[r]+=n; st[r]+=s; ops[r][txt.split()[0] if not txt.strip().startswith("@") else txt.split()[1]]+=n
tot=sum(cnt.values())+sum(spin.values()); ts=sum(st.values())
print("total", tot, "per tile", tot/5970)
for r,n in sorted(cnt.items(), key=lambda kv:-kv[1]):
    print(f"{r:16s} {n:11d} {100*n/tot:5.1f}%  per-tile {n/5970:8.0f}  spin {spin[r]:10d}  samples work {100*st[r]/ts:5.1f}% wait {100*st[r+'/wait']/ts:5.1f}%")
    if len(sys.argv)>4: print("      ", ", ".join(f"{k}:{v/5970:.0f}" for k,v in ops[r].most_common(14)))
