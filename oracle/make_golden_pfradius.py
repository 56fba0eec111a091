"""Generate tests/golden/pf_radius.npz: one denoiser call of the reference's OWN PharmRecDynamicsGVP with pf_k = 0 -- pf / fp
edges from radius(pharm, prot, r = graph_cutoffs['pf'], max_num_neighbors = 100) instead of kNN (dynamics_gvp.py:210-216; the
constructor default of the reference, configs/dev.yml uses pf_k = 5) -- on a 400-atom pocket, where a pharmacophore centre
collects far more than 128 in-edges.  CPU, over the pure-torch shims.  Test infrastructure only.

    python oracle/make_golden_pfradius.py
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    MG.reference_loader.load()
    import yaml
    from pharmacoforge.config_utils.load_from_config import model_from_config
    from pharmacoforge.utils import get_batch_idxs
    from pharmacoforge_b200.synthetic import synth_state_dict
    cfg = copy.deepcopy(yaml.safe_load(open(os.path.join(MG.reference_loader.REFERENCE_ROOT, "configs", "dev.yml"))))
    cfg["dynamics"]["pf_k"] = 0
    cfg["graph"]["graph_cutoffs"]["pf"] = 11.0      # (dev.yml: 8) wide enough for in-degrees well above one 128-row tile
    torch.manual_seed(0)
    model = model_from_config(cfg)
    layout = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = synth_state_dict(layout, seed=0)
    sd["gamma.gamma"] = model.state_dict()["gamma.gamma"]
    model.load_state_dict(sd, strict=True)
    model.eval()
    sizes = [3, 8, 5]
    gb, _ = MG.make_batch(cfg, pocket_seed=2, n_atoms=400, sizes=sizes)
    gen = torch.Generator().manual_seed(91)
    nf = sum(sizes)
    x_t = torch.randn(nf, 3, generator=gen) * 4.0
    x_t[0] = torch.tensor([40.0, 0.0, 0.0])          # one centre far outside the pocket: no pf edges at all
    h_t = torch.randn(nf, 6, generator=gen)
    tt = torch.tensor([0.2, 0.8, 0.55])
    bi = get_batch_idxs(gb)
    prot_shift = torch.randn(len(sizes), 3, generator=gen) * 1.5
    gb.nodes["prot"].data["x_0"] = gb.nodes["prot"].data["x_0"] - gb.nodes["prot"].data["x_0"].mean(0, keepdim=True) \
        + prot_shift[bi["prot"]]
    gb.nodes["pharm"].data["x_t"] = x_t
    gb.nodes["pharm"].data["h_t"] = h_t
    cap = {}
    dyn = model.dynamics

    def grab_edges(mod, args):
        g = args[0]
        for et in ("ff", "pf", "fp"):
            a, b = g.edges(form="uv", etype=et)
            cap["e_" + et] = MG.sorted_edges(a, b)
    hooks = [dyn.noise_predictor.register_forward_pre_hook(grab_edges)]
    conv = dyn.noise_predictor.conv_layers[1]

    def grab(mod, args, out):
        cap["conv1_pharm_h"] = out["pharm"][0].detach().clone()
        cap["conv1_prot_h"] = out["prot"][0].detach().clone()
    hooks.append(conv.register_forward_hook(grab))
    with torch.no_grad():
        prot_x_in = gb.nodes["prot"].data["x_0"].clone()
        eps_h, eps_x = dyn(gb, tt, bi)
    for h in hooks:
        h.remove()
    out = dict(sizes=np.array(sizes, np.int32), n_atoms=np.int32(400), pocket_seed=np.int32(2), prot_x=prot_x_in.numpy(),
               x_t=x_t.numpy(), h_t=h_t.numpy(), t=tt.numpy(), eps_h=eps_h.numpy(), eps_x=eps_x.numpy(),
               pf_cutoff=np.float64(cfg["graph"]["graph_cutoffs"]["pf"]))
    for k, val in cap.items():
        if k.startswith("e_"):
            out[k + "_src"], out[k + "_dst"] = val
        else:
            out[k] = val.numpy()
    deg = np.bincount(out["e_pf_dst"], minlength=nf)
    print("pf edges", out["e_pf_src"].shape[0], "in-degree per pharm node", deg.tolist())
    np.savez_compressed(os.path.join(MG.GOLD, "pf_radius.npz"), **out)
    print(os.path.getsize(os.path.join(MG.GOLD, "pf_radius.npz")))


if __name__ == "__main__":
    main()
