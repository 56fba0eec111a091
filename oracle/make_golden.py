"""Generate tests/golden/* by running the reference's OWN model code on CPU.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference).
The reference's pharmacoforge/models/{pharmacodiff,dynamics_gvp,gvp}.py,
utils/unorganized_utils.py and dataset/protein_pharm_dataset.py are imported
unmodified over oracle/shims (pure-torch stand-ins for dgl / torch_cluster /
pytorch_lightning, none of which is installable here).  Nothing from the
reference is copied: only its numerical outputs on seeded inputs are stored.

    python oracle/make_golden.py          # rewrites tests/golden/
"""
import json
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import reference_loader  # noqa: E402

from pharmacoforge_b200.synthetic import make_pocket, synth_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_model():
    reference_loader.load()
    from pharmacoforge.config_utils.load_from_config import model_from_config

    cfg = yaml.safe_load(open(os.path.join(reference_loader.REFERENCE_ROOT, "configs", "dev.yml")))
    torch.manual_seed(0)
    model = model_from_config(cfg)
    layout = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = synth_state_dict(layout, seed=0)
    sd["gamma.gamma"] = model.state_dict()["gamma.gamma"]
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, cfg, layout


def make_batch(cfg, pocket_seed, n_atoms, sizes):
    import dgl
    from pharmacoforge.dataset.protein_pharm_dataset import build_initial_complex_graph
    from pharmacoforge.utils import copy_graph

    pos, onehot = make_pocket(n_atoms, seed=pocket_seed)
    ref_graph = build_initial_complex_graph(torch.from_numpy(pos), torch.from_numpy(onehot),
                                            cutoffs=cfg["graph"]["graph_cutoffs"])
    copies = copy_graph(ref_graph, len(sizes), pharm_feats_per_copy=torch.tensor(sizes))
    return dgl.batch(copies), ref_graph


def sorted_edges(u, v):
    """canonical (dst, src) order, int32"""
    u = u.numpy().astype(np.int64)
    v = v.numpy().astype(np.int64)
    order = np.lexsort((u, v))
    return u[order].astype(np.int32), v[order].astype(np.int32)


class InjectedRandn:
    """Replaces torch.randn inside the reference's sampler with rows of a pre-drawn buffer, consumed in
    the reference's call order (x then h: pharmacodiff.py:455-456, :423-424)."""

    def __init__(self, noise):
        self.noise = noise
        self.calls = 0

    def __call__(self, *shape, **kwargs):
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        row, is_h = divmod(self.calls, 2)
        self.calls += 1
        out = self.noise[row, :, 3:9] if is_h else self.noise[row, :, 0:3]
        assert tuple(out.shape) == tuple(shape), (out.shape, shape)
        return out.clone()


def main():
    os.makedirs(GOLD, exist_ok=True)
    model, cfg, layout = build_model()
    from pharmacoforge.models.gvp import _norm_no_nan, _rbf
    from pharmacoforge.utils import get_batch_idxs

    with open(os.path.join(GOLD, "state_dict_layout.json"), "w") as f:
        json.dump(layout, f, indent=0, sort_keys=True)

    # ---- (1) known-answer constants: schedule + posterior coefficients, straight from the reference
    T = model.n_timesteps
    s = torch.arange(T).float() / T
    t = (torch.arange(T) + 1).float() / T
    g_s, g_t = model.gamma(s), model.gamma(t)
    sigma2_ts, sigma_ts, alpha_ts, alpha_s = model.sigma_and_alpha_t_given_s(g_t, g_s)
    sigma_s, sigma_t = model.sigma(g_s), model.sigma(g_t)
    var_terms = sigma2_ts / alpha_ts / sigma_t
    sigma_q = sigma_ts * sigma_s / sigma_t
    np.savez(os.path.join(GOLD, "constants.npz"),
             gamma=model.gamma.gamma.detach().numpy(), alpha_ts=alpha_ts.detach().numpy(),
             var_terms=var_terms.detach().numpy(), sigma_q=sigma_q.detach().numpy(),
             sigma_t=sigma_t.detach().numpy(), alpha_t=model.alpha(g_t).detach().numpy(),
             rbf3=_rbf(torch.tensor([3.0]), D_max=15, D_count=16).numpy(),
             rbf_grid=_rbf(torch.linspace(0, 20, 41), D_max=15, D_count=16).numpy(),
             norm0=_norm_no_nan(torch.zeros(1, 3)).numpy())

    # ---- (2) static pp radius graph of the config-1 pocket
    gb, ref_graph = make_batch(cfg, pocket_seed=0, n_atoms=400, sizes=[3])
    u, v = ref_graph.edges(form="uv", etype="pp")
    su, sv = sorted_edges(u, v)
    np.savez_compressed(os.path.join(GOLD, "pp_graph_n400_seed0.npz"), src=su, dst=sv)

    # ---- (3) one teacher-forced denoiser call with per-kernel intermediates
    sizes = [3, 5, 8, 6]
    gb, _ = make_batch(cfg, pocket_seed=3, n_atoms=100, sizes=sizes)
    gen = torch.Generator().manual_seed(77)
    nf = sum(sizes)
    x_t = torch.randn(nf, 3, generator=gen) * 3.0
    h_t = torch.randn(nf, 6, generator=gen)
    tt = torch.tensor([0.37, 0.99, 0.01, 0.5])
    # put every graph in its own pharmacophore-COM frame, as the sampler does each step
    prot_shift = torch.randn(len(sizes), 3, generator=gen) * 2.0
    bi = get_batch_idxs(gb)
    gb.nodes["prot"].data["x_0"] = gb.nodes["prot"].data["x_0"] - gb.nodes["prot"].data["x_0"].mean(0, keepdim=True) \
        + prot_shift[bi["prot"]]
    gb.nodes["pharm"].data["x_t"] = x_t
    gb.nodes["pharm"].data["h_t"] = h_t
    cap = {}
    dyn = model.dynamics
    hooks = [
        dyn.pharm_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("enc_pharm", o.detach().clone())),
        dyn.prot_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("enc_prot", o.detach().clone())),
    ]

    def grab_edges(mod, args):
        g = args[0]
        for et in ("ff", "pf", "fp", "pp"):
            a, b = g.edges(form="uv", etype=et)
            cap["e_" + et] = sorted_edges(a, b)

    hooks.append(dyn.noise_predictor.register_forward_pre_hook(grab_edges))
    for li, conv in enumerate(dyn.noise_predictor.conv_layers):
        def grab(mod, args, out, li=li):
            for nt in ("pharm", "prot"):
                cap[f"conv{li}_{nt}_h"] = out[nt][0].detach().clone()
                cap[f"conv{li}_{nt}_v"] = out[nt][2].detach().clone()
        hooks.append(conv.register_forward_hook(grab))
    with torch.no_grad():
        prot_x_in = gb.nodes["prot"].data["x_0"].clone()
        eps_h, eps_x = dyn(gb, tt, bi)
    for h in hooks:
        h.remove()
    out = dict(sizes=np.array(sizes, np.int32), n_atoms=np.int32(100), pocket_seed=np.int32(3),
               prot_x=prot_x_in.numpy(), x_t=x_t.numpy(), h_t=h_t.numpy(), t=tt.numpy(),
               eps_h=eps_h.numpy(), eps_x=eps_x.numpy())
    for k, val in cap.items():
        if k.startswith("e_"):
            out[k + "_src"], out[k + "_dst"] = val
        else:
            out[k] = val.numpy()
    np.savez_compressed(os.path.join(GOLD, "denoiser_call.npz"), **out)

    # ---- (4) a full T-step reverse diffusion with injected noise (trajectory for teacher forcing)
    sizes = [4, 7]
    gb, _ = make_batch(cfg, pocket_seed=5, n_atoms=100, sizes=sizes)
    nf = sum(sizes)
    noise = torch.randn(T + 1, nf, 9, generator=torch.Generator().manual_seed(1234))
    traj = []
    orig = model.sample_p_zs_given_zt

    def traced(s_arr, t_arr, g, batch_idxs):
        if not traj:
            traj.append((g.nodes["pharm"].data["x_t"].clone(), g.nodes["pharm"].data["h_t"].clone(),
                         g.nodes["prot"].data["x_0"].clone()))
        g = orig(s_arr, t_arr, g, batch_idxs)
        traj.append((g.nodes["pharm"].data["x_t"].clone(), g.nodes["pharm"].data["h_t"].clone(),
                     g.nodes["prot"].data["x_0"].clone()))
        return g

    model.sample_p_zs_given_zt = traced
    real_randn = torch.randn
    torch.randn = InjectedRandn(noise)
    try:
        pharms = model.sample_given_receptor(gb)
    finally:
        torch.randn = real_randn
        model.sample_p_zs_given_zt = orig
    assert len(traj) == T + 1
    np.savez_compressed(
        os.path.join(GOLD, "sample_traj.npz"), sizes=np.array(sizes, np.int32), n_atoms=np.int32(100),
        pocket_seed=np.int32(5), noise=noise.numpy(),
        traj_x=torch.stack([a for a, _, _ in traj]).numpy(), traj_h=torch.stack([b for _, b, _ in traj]).numpy(),
        traj_prot0=torch.stack([c[[0, 100]] for _, _, c in traj]).numpy(),
        final_x=torch.cat([p.ph_coords for p in pharms]).numpy(),
        final_h=torch.cat([p.g.nodes["pharm"].data["h_0"] for p in pharms]).numpy(),
        final_type=torch.cat([p.ph_feats_idxs for p in pharms]).numpy().astype(np.int32),
        final_prot=torch.cat([p.g.nodes["prot"].data["x_0"] for p in pharms]).numpy())
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    main()
