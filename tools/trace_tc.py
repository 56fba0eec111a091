import os, sys, json, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmacoforge_b200 import ops, _lib
from pharmacoforge_b200.batch import GraphBatch, Pocket
from pharmacoforge_b200.diffusion import PharmacophoreDiff
from pharmacoforge_b200.hostutil import polynomial_gamma
from pharmacoforge_b200.synthetic import make_pocket, synth_state_dict, readme_sizes
layout = json.load(open(os.path.join(ROOT, "tests/golden/state_dict_layout.json")))
sd = synth_state_dict(layout, seed=0); sd["gamma.gamma"] = polynomial_gamma(100, 1e-5, 2.0)
dyn = dict(vector_size=16, n_convs=2, n_hidden_scalars=128, message_norm="mean", dropout=0.1, ff_k=0, pf_k=5, n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4)
cut = {"pp": 3.5, "pf": 8, "fp": 8, "ff": 9}
model = PharmacophoreDiff(6, 11, ["a","b","c","d","e","f"], n_timesteps=100, graph_config={"graph_cutoffs": cut}, dynamics_config=dyn, precision=1e-5)
model.load_state_dict(sd); model.eval()
dev = torch.device("cuda:0")
npk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
FP16 = len(sys.argv) > 3 and sys.argv[3] == 'fp16'
pockets = [Pocket.from_numpy(*make_pocket(400, seed=i)) for i in range(npk)]
g = GraphBatch.from_pockets(pockets, [readme_sizes(30)] * npk, dev)
W = model.dynamics.packed_weights(dev)
lib = _lib.load()
prot_h = torch.randn(g.n_prot, 128, device=dev); prot_v = torch.randn(g.n_prot, 48, device=dev)
agg_h = torch.zeros(g.n_prot, 128, device=dev); agg_v = torch.zeros(g.n_prot, 48, device=dev)
blob = W.tc[3 * W.tc_stride:4 * W.tc_stride]
def run(v):
    ops.edge_conv_tc(prot_h, v, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles, g.pp_n_tiles, blob, agg_h, agg_v, False, FP16)
for v in (None, prot_v):
    run(v); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(v); e1.record(); torch.cuda.synchronize()
    print("HAS_V", v is not None, "edges", g.n_pp_edges, "tiles", int(g.pp_n_tiles), "ms", e0.elapsed_time(e1))
MODE = sys.argv[4] if len(sys.argv) > 4 else "l1"      # l1: general kernel with vectors; seed: seeded first-layer kernel
seed_row, seed_rep = g.seed_arrays()
table = torch.randn(seed_rep.numel(), 128, device=dev)
blob0 = W.tc[3 * W.tc_stride:4 * W.tc_stride]
def run_seed():
    ops.edge_conv_tc_seeded(seed_row, table, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles, g.pp_n_tiles, blob0, agg_h, agg_v, False, FP16)
run_seed(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run_seed(); e1.record(); torch.cuda.synchronize()
print("SEEDED edges", g.n_pp_edges, "ms", e0.elapsed_time(e1))
tr = torch.zeros((4 * 4096 + 148) * 2, dtype=torch.int64, device=dev)
lib.pf_tc_trace(C.c_void_p(tr.data_ptr()))
if MODE == "seed":
    run_seed()
else:
    run(prot_v)
torch.cuda.synchronize()
lib.pf_tc_trace(None)
cta = tr.cpu().numpy()[4 * 4096 * 2:].reshape(148, 2)
t = tr.cpu().numpy()[:4 * 4096 * 2].reshape(4, 4096, 2)
b0 = cta[:, 0].min()
print('CTA spans us: begin', np.round((cta[:, 0] - b0) / 1e3, 1).tolist()[:16], 'dur min/med/max', float(np.min(cta[:,1]-cta[:,0]))/1e3, float(np.median(cta[:,1]-cta[:,0]))/1e3, float(np.max(cta[:,1]-cta[:,0]))/1e3, 'kernel', float(cta[:,1].max()-b0)/1e3)
print('dur per CTA us', np.round((cta[:,1]-cta[:,0])/1e3).astype(int).tolist())
t0 = min(t[r, 0, 1] for r in range(4) if t[r, 0, 1] > 0)
names = {0x01: "tile start", 0x02: "meta done", 0x03: "gather done", 0x10: "EPI-A start", 0x11: "vecD ok", 0x12: "A arrived", 0x21: "D ok", 0x22: "F arrived", 0x31: "gate ok", 0x32: "vecA arrived"}
mn = {0x10: "V issue", 0x11: "V committed", 0x20: "S start", 0x21: "S committed", 0x30: "G issue", 0x31: "G committed"}
ev = []
for r in range(4):
    for i in range(4096):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0: break
        if r < 2:
            ev.append((clk - t0, f"slot{r} epi g={tag >> 8} {names.get(tag & 0xff, hex(tag & 0xff))}"))
        else:
            if 0x40 <= (tag & 0xff) < 0x50: continue
            ev.append((clk - t0, f"      MMA slot{r - 2} g={(tag >> 8) & 15} {mn.get(tag & 0xff, hex(tag & 0xff))}"))
ev.sort()
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 130
prev = 0
for c, s in ev[:lim]:
    print(f"{c:9d} (+{c - prev:6d}) {s}")
    prev = c
