import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pf_oracle as O
from pharmacoforge_b200.batch import GraphBatch, Pocket
from pharmacoforge_b200.diffusion import PharmacophoreDiff
from pharmacoforge_b200.synthetic import make_pocket, synth_state_dict
layout = json.load(open(os.path.join(ROOT, "tests/golden/state_dict_layout.json")))
sd = synth_state_dict(layout, seed=0); sd["gamma.gamma"] = O.gamma_table(100, 1e-5)
dyn = dict(vector_size=16, n_convs=2, n_hidden_scalars=128, message_norm="mean", dropout=0.1, ff_k=0, pf_k=5, n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4)
cut = {"pp": 3.5, "pf": 8, "fp": 8, "ff": 9}
model = PharmacophoreDiff(6, 11, ["a","b","c","d","e","f"], n_timesteps=100, graph_config={"graph_cutoffs": cut}, dynamics_config=dyn, precision=1e-5)
model.load_state_dict(sd); model.eval()
d = dict(np.load(os.path.join(ROOT, "tests/golden/denoiser_call.npz")))
t = lambda a: torch.from_numpy(np.asarray(a))
def from_cm(v): return v.reshape(v.shape[0], 3, 16).permute(0, 2, 1).contiguous()
sizes = [int(v) for v in d["sizes"]]
pos, onehot = make_pocket(int(d["n_atoms"]), seed=int(d["pocket_seed"]))
res = {}
for tr in (64, 128):
    g = GraphBatch.from_pockets([Pocket.from_numpy(pos, onehot)], [sizes], "cuda:0", tile_rows=tr)
    st = model.dynamics.bind(g)
    g.pharm_x.copy_(t(d["x_t"]).cuda()); g.pharm_h.copy_(t(d["h_t"]).cuda()); g.prot_x.copy_(t(d["prot_x"]).cuda())
    eps_h, eps_x = model.dynamics(g, t(d["t"]), None)
    torch.cuda.synchronize(); g.check_status()
    out = {"conv1_pharm_h": st.pharm_hh, "conv1_pharm_v": from_cm(st.pharm_v), "conv1_prot_h": st.prot_h, "conv1_prot_v": from_cm(st.prot_v), "eps_h": eps_h, "eps_x": eps_x}
    res[tr] = {k: v.detach().cpu().double().numpy().copy() for k, v in out.items()}
    for k, v in res[tr].items():
        ref = d[k].astype(np.float64); err = np.abs(v - ref); bound = 1e-5 + 1e-4 * np.abs(ref)
        i = np.unravel_index(np.argmax(err / bound), err.shape)
        print(f"tile_rows={tr} {k:14s} max abs err {err.max():.3e} worst ratio {(err/bound).max():.2f} at {i} ref {ref[i]:.4e} got {v[i]:.4e}  rms ref {np.sqrt((ref**2).mean()):.3e}")
for k in res[64]:
    print("tc vs ffma", k, np.abs(res[64][k] - res[128][k]).max())
