"""K3 / K4 launch times at bench scale, outside the sampling loop.  usage: kbench.py <pockets> [reps] [fp16]
Prints one JSON line: avg ms of the pp edge conv (layer 0: no source vectors, layer 1: with) and of the protein
node update, with L2-exceeding inputs at >= 64 pockets (2.2 GB of node rows at 256)."""
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
cut = {"pp": 3.5, "pf": 8, "fp": 8, "ff": 9}
model = PharmacophoreDiff(6, 11, list("abcdef"), n_timesteps=100, graph_config={"graph_cutoffs": cut}, dynamics_config=dyn, precision=1e-5)
model.load_state_dict(sd); model.eval()
dev = torch.device("cuda:0")
npk = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
FP16 = len(sys.argv) > 3 and sys.argv[3] == "fp16"
pockets = [Pocket.from_numpy(*make_pocket(400, seed=i)) for i in range(npk)]
g = GraphBatch.from_pockets(pockets, [readme_sizes(30)] * npk, dev)
W = model.dynamics.packed_weights(dev)
torch.manual_seed(0)
prot_h = torch.randn(g.n_prot, 128, device=dev); prot_v = 0.1 * torch.randn(g.n_prot, 48, device=dev)
agg_h = torch.zeros(g.n_prot, 128, device=dev); agg_v = torch.zeros(g.n_prot, 48, device=dev)
out_h = torch.empty_like(prot_h); out_v = torch.empty_like(prot_v)


def timed(fn):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def k3(layer, v):
    blob = W.tc[(4 * layer + 3) * W.tc_stride:(4 * layer + 4) * W.tc_stride]
    return lambda: ops.edge_conv_tc(prot_h, v, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles,
                                    g.pp_n_tiles, blob, agg_h, agg_v, False, FP16)


def k4(layer, v):
    return lambda: ops.node_update_tc(prot_h, v, agg_h, agg_v, W.tcu_view(layer, 1), out_h, out_v, FP16)


only = os.environ.get("KB_ONLY", "")
res = {"pockets": npk, "edges": g.n_pp_edges, "nodes": g.n_prot, "fp16": FP16, "lib": os.environ.get("PF_LIB_PATH", "")}
if only in ("", "k3"):
    res.update(k3_l0_ms=timed(k3(0, None)), k3_l1_ms=timed(k3(1, prot_v)))
    res["k3_tflops_l1"] = g.n_pp_edges * 136742 / (res["k3_l1_ms"] * 1e-3) / 1e12
if only in ("", "k4"):
    res.update(k4_l0_ms=timed(k4(0, None)), k4_l1_ms=timed(k4(1, prot_v)))
    res["k4_gbs_l1"] = 3 * 704 * g.n_prot / (res["k4_l1_ms"] * 1e-3) / 1e9
    # correctness guard for A/B builds: checksum of the outputs
    res["k4_checksum"] = float(out_h.double().sum().item()), float(out_v.double().sum().item())
print(json.dumps(res))
