"""Device timeline of CTA 0 of the tcgen05 node-update kernel (K4).  usage: trace_k4.py <pockets> <events> [fp16]"""
import os, sys, json, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmacoforge_b200 import ops, _lib
from pharmacoforge_b200.diffusion import PharmacophoreDiff
from pharmacoforge_b200.hostutil import polynomial_gamma
from pharmacoforge_b200.synthetic import synth_state_dict
layout = json.load(open(os.path.join(ROOT, "tests/golden/state_dict_layout.json")))
sd = synth_state_dict(layout, seed=0); sd["gamma.gamma"] = polynomial_gamma(100, 1e-5, 2.0)
dyn = dict(vector_size=16, n_convs=2, n_hidden_scalars=128, message_norm="mean", dropout=0.1, ff_k=0, pf_k=5, n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4)
cut = {"pp": 3.5, "pf": 8, "fp": 8, "ff": 9}
model = PharmacophoreDiff(6, 11, list("abcdef"), n_timesteps=100, graph_config={"graph_cutoffs": cut}, dynamics_config=dyn, precision=1e-5)
model.load_state_dict(sd); model.eval()
dev = torch.device("cuda:0")
npk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
FP16 = len(sys.argv) > 3 and sys.argv[3] == "fp16"
n = npk * 30 * 400
W = model.dynamics.packed_weights(dev)
lib = _lib.load()
h = torch.randn(n, 128, device=dev); v = torch.randn(n, 48, device=dev)
ah = torch.randn(n, 128, device=dev); av = torch.randn(n, 48, device=dev)
def run():
    ops.node_update_tc(h, v, ah, av, W.tcu_view(1, 1), h, v, FP16)
for _ in range(2):
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print("nodes", n, "tiles", (n + 127) // 128, "ms", e0.elapsed_time(e1))
tr = torch.zeros((4 * 4096 + 148) * 2, dtype=torch.int64, device=dev)
lib.pf_tc_trace(C.c_void_p(tr.data_ptr()))
run(); torch.cuda.synchronize()
lib.pf_tc_trace(None)
t = tr.cpu().numpy()[:4 * 4096 * 2].reshape(4, 4096, 2)
t0 = min(t[r, 0, 1] for r in range(4) if t[r, 0, 1] > 0)
names = {0x01: "tile start", 0x41: "front H done", 0x42: "front H barrier", 0x10: "vec staged / GVP start", 0x43: "GVPs done", 0x44: "back V done", 0x45: "back H stats done", 0x46: "V loads stored", 0x47: "V loads barrier", 0x48: "V normalised", 0x49: "V barrier 2", 0x4a: "front H loads arrived", 0x4b: "front H stats exchanged"}
mn = {0x10: "V issue", 0x11: "V committed", 0x20: "S start", 0x21: "S committed", 0x30: "G issue", 0x31: "G committed"}
ev = []
for r in range(4):
    for i in range(4096):
        tag, clk = int(t[r, i, 0]), int(t[r, i, 1])
        if clk == 0: break
        if r < 2: ev.append((clk - t0, f"slot{r} epi g={tag >> 8} {names.get(tag & 0xff, hex(tag & 0xff))}"))
        else: ev.append((clk - t0, f"      MMA slot{r - 2} g={(tag >> 8) & 15} {mn.get(tag & 0xff, hex(tag & 0xff))}"))
ev.sort()
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 130
prev = 0
for c, s in ev[:lim]:
    print(f"{c:9d} (+{c - prev:6d}) {s}"); prev = c
