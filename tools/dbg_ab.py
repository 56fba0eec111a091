import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pf_oracle as O
from pharmacoforge_b200 import ops
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
sizes = [int(v) for v in d["sizes"]]
pos, onehot = make_pocket(int(d["n_atoms"]), seed=int(d["pocket_seed"]))
dev = "cuda:0"
W = model.dynamics.packed_weights(torch.device(dev))
out = {}
for tr in (64, 128):
    g = GraphBatch.from_pockets([Pocket.from_numpy(pos, onehot)], [sizes], dev, tile_rows=tr)
    st = model.dynamics.bind(g)
    g.pharm_x.copy_(t(d["x_t"]).cuda()); g.pharm_h.copy_(t(d["h_t"]).cuda()); g.prot_x.copy_(t(d["prot_x"]).cuda())
    tt = t(d["t"]).float().cuda()
    prot_h = ops.encode(g.prot_feats, g.prot_ptr, tt, W.view("prot_enc"))
    agg_h = torch.zeros(g.n_prot, 128, device=dev); agg_v = torch.zeros(g.n_prot, 48, device=dev)
    if tr == 128:
        blob = W.tc[3 * W.tc_stride:4 * W.tc_stride]
        ops.edge_conv_tc(prot_h, None, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles, g.pp_n_tiles, blob, agg_h, agg_v, False)
    else:
        ops.edge_conv(prot_h, None, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles, g.pp_n_tiles, W.view("msg0_3"), 3, agg_h, agg_v, False)
    torch.cuda.synchronize()
    out[tr] = (agg_h.cpu().double(), agg_v.cpu().double())
# oracle messages in float64 for truth
b = O.build_batch([(t(pos), t(onehot))], [sizes])
b.prot_x = t(d["prot_x"]).clone()
for name, i in (("agg_h", 0), ("agg_v", 1)):
    a, c = out[64][i], out[128][i]
    diff = (a - c).abs()
    print(name, "ffma rms", float(a.pow(2).mean().sqrt()), "max|tc-ffma|", float(diff.max()), "rel to rms", float(diff.max() / a.pow(2).mean().sqrt()))
    rowerr = diff.max(dim=1).values
    print("  rows with err > 10x median:", int((rowerr > 10 * rowerr.median()).sum()), "median row err", float(rowerr.median()), "max row", int(rowerr.argmax()))
    r = int(rowerr.argmax())
    print("  row", r, "deg", int(g.pp_cnt[r]), "ffma", a[r, :6].numpy(), "tc", c[r, :6].numpy())
