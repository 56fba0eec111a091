"""Launch target for ncu: the bench-scale pp edge conv (seeded first-layer kernel, then the general kernel with source
vectors) and the protein node update, each once after one warm-up launch.  usage: ncu_target.py <pockets>"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmacoforge_b200 import ops
from pharmacoforge_b200.batch import GraphBatch, Pocket
from pharmacoforge_b200.diffusion import PharmacophoreDiff
from pharmacoforge_b200.hostutil import polynomial_gamma
from pharmacoforge_b200.synthetic import make_pocket, synth_state_dict, readme_sizes
layout = json.load(open(os.path.join(ROOT, "tests/golden/state_dict_layout.json")))
sd = synth_state_dict(layout, seed=0); sd["gamma.gamma"] = polynomial_gamma(100, 1e-5, 2.0)
dyn = dict(vector_size=16, n_convs=2, n_hidden_scalars=128, message_norm="mean", dropout=0.1, ff_k=0, pf_k=5, n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4)
model = PharmacophoreDiff(6, 11, list("abcdef"), n_timesteps=100, graph_config={"graph_cutoffs": {"pp": 3.5, "pf": 8, "fp": 8, "ff": 9}}, dynamics_config=dyn, precision=1e-5)
model.load_state_dict(sd); model.eval()
dev = torch.device("cuda:0")
npk = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = GraphBatch.from_pockets([Pocket.from_numpy(*make_pocket(400, seed=i)) for i in range(npk)], [readme_sizes(30)] * npk, dev)
W = model.dynamics.packed_weights(dev)
torch.manual_seed(0)
prot_h = torch.randn(g.n_prot, 128, device=dev); prot_v = 0.1 * torch.randn(g.n_prot, 48, device=dev)
agg_h = torch.zeros(g.n_prot, 128, device=dev); agg_v = torch.zeros(g.n_prot, 48, device=dev)
seed_row, seed_rep = g.seed_arrays()
table = torch.randn(seed_rep.numel(), 128, device=dev)
b0, b1 = W.tc[3 * W.tc_stride:4 * W.tc_stride], W.tc[7 * W.tc_stride:8 * W.tc_stride]
for _ in range(2):   # launch order per pass: seeded K3, general K3 (layer 1), K4 (layer 1)
    ops.edge_conv_tc_seeded(seed_row, table, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles, g.pp_n_tiles, b0, agg_h, agg_v, False, False)
    ops.edge_conv_tc(prot_h, prot_v, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles, g.pp_n_tiles, b1, agg_h, agg_v, False, False)
    ops.node_update_tc(prot_h, prot_v, agg_h, agg_v, W.tcu_view(1, 1), prot_h, prot_v, False)
    torch.cuda.synchronize()
print("edges", g.n_pp_edges, "nodes", g.n_prot)
