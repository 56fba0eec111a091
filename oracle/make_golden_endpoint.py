"""Generate tests/golden/endpoint_param.npz: the reference's OWN PharmacophoreDiff with the endpoint parameterisation
(`endpoint_param_feat = endpoint_param_coord = True`, pharmacodiff.py:204-216, 413-420 -- the mode README.md's
`configs/endpoint_param.yaml` training command refers to) and with `remove_com = False` (pharmacodiff.py:123-125),
run on CPU over the pure-torch shims with injected timesteps / Gaussian draws.  Test infrastructure only.

    python oracle/make_golden_endpoint.py
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def build(flags):
    MG.reference_loader.load()
    import yaml
    from pharmacoforge.config_utils.load_from_config import model_from_config
    from pharmacoforge_b200.synthetic import synth_state_dict
    cfg = yaml.safe_load(open(os.path.join(MG.reference_loader.REFERENCE_ROOT, "configs", "dev.yml")))
    cfg = copy.deepcopy(cfg)
    cfg["diffusion"].update(flags)
    torch.manual_seed(0)
    model = model_from_config(cfg)
    layout = {k: list(v.shape) for k, v in model.state_dict().items()}
    sd = synth_state_dict(layout, seed=0)
    sd["gamma.gamma"] = model.state_dict()["gamma.gamma"]
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, cfg


def run_forward(model, cfg, sizes, x0, h0, t_int, eps_h, eps_x):
    gb, _ = MG.make_batch(cfg, pocket_seed=5, n_atoms=120, sizes=sizes)
    gb.nodes["pharm"].data["x_0"] = x0.clone()
    gb.nodes["pharm"].data["h_0"] = h0.clone()
    draws = [eps_h, eps_x]
    real_randn, real_randint = torch.randn, torch.randint

    def fake_randn(*shape, **kw):
        return draws.pop(0).clone()

    def fake_randint(low, high, size, **kw):
        return t_int.clone()

    torch.randn, torch.randint = fake_randn, fake_randint
    try:
        losses, metrics = model.forward(gb, phase="val")
    finally:
        torch.randn, torch.randint = real_randn, real_randint
    total = torch.stack(list(losses.values())).sum()
    model.zero_grad()
    total.backward()
    names, norms = [], []
    for k, p_ in sorted(model.named_parameters()):
        if p_.numel() and p_.grad is not None:
            names.append(k)
            norms.append(float(p_.grad.double().norm()))
    out = {k.replace(" ", "_"): float(v) for k, v in {**losses, **metrics}.items()}
    return out, names, norms


def main():
    sizes = [4, 6, 8, 5]
    nf = sum(sizes)
    gen = torch.Generator().manual_seed(4242)
    save = {}
    for tag, flags in (("ep", dict(endpoint_param_feat=True, endpoint_param_coord=True)),
                       ("epx", dict(endpoint_param_feat=False, endpoint_param_coord=True)),
                       ("nocom", dict(remove_com=False)),
                       ("ep_nocom", dict(endpoint_param_feat=True, endpoint_param_coord=True, remove_com=False))):
        model, cfg = build(flags)
        if "x0" not in save:
            gb, _ = MG.make_batch(cfg, pocket_seed=5, n_atoms=120, sizes=sizes)
            centre = gb.nodes["prot"].data["x_0"].mean(0, keepdim=True)
            save["x0"] = (centre + torch.randn(nf, 3, generator=gen) * 2.5).numpy()
            save["h0"] = torch.nn.functional.one_hot(torch.randint(0, 6, (nf,), generator=gen), 6).float().numpy()
            save["t_int"] = np.asarray([37, 99, 0, 63])
            save["eps_h"] = torch.randn(nf, 6, generator=gen).numpy()
            save["eps_x"] = torch.randn(nf, 3, generator=gen).numpy()
        t = lambda k: torch.from_numpy(save[k])
        out, names, norms = run_forward(model, cfg, sizes, t("x0"), t("h0"), t("t_int"), t("eps_h"), t("eps_x"))
        print(tag, out, len(names))
        for k, v in out.items():
            save[f"{tag}__{k}"] = np.float64(v)
        save[f"{tag}__grad_names"] = np.asarray(names)
        save[f"{tag}__grad_norms"] = np.asarray(norms)

    # ---- reverse diffusion with the endpoint posterior (pharmacodiff.py:413-420), full T steps, injected noise
    for tag, flags in (("ep", dict(endpoint_param_feat=True, endpoint_param_coord=True)),
                       ("eph", dict(endpoint_param_feat=True, endpoint_param_coord=False))):
        model, cfg = build(flags)
        T = model.n_timesteps
        ssz = [4, 7]
        gb, _ = MG.make_batch(cfg, pocket_seed=5, n_atoms=100, sizes=ssz)
        noise = torch.randn(T + 1, sum(ssz), 9, generator=torch.Generator().manual_seed(99))
        traj = []
        orig = model.sample_p_zs_given_zt

        def traced(s_arr, t_arr, g, batch_idxs):
            if not traj:
                traj.append((g.nodes["pharm"].data["x_t"].clone(), g.nodes["pharm"].data["h_t"].clone()))
            g = orig(s_arr, t_arr, g, batch_idxs)
            traj.append((g.nodes["pharm"].data["x_t"].clone(), g.nodes["pharm"].data["h_t"].clone()))
            return g

        model.sample_p_zs_given_zt = traced
        real_randn = torch.randn
        torch.randn = MG.InjectedRandn(noise)
        try:
            with torch.no_grad():
                pharms = model.sample_given_receptor(gb)
        finally:
            torch.randn = real_randn
            model.sample_p_zs_given_zt = orig
        save[f"s_{tag}__noise"] = noise.numpy()
        save[f"s_{tag}__traj_x"] = torch.stack([a for a, _ in traj]).numpy()
        save[f"s_{tag}__traj_h"] = torch.stack([b for _, b in traj]).numpy()
        save[f"s_{tag}__final_x"] = torch.cat([p.ph_coords for p in pharms]).numpy()
        save[f"s_{tag}__final_h"] = torch.cat([p.g.nodes["pharm"].data["h_0"] for p in pharms]).numpy()
        print("sample", tag, save[f"s_{tag}__final_x"][:2])
    # posterior coefficients of the endpoint branch, straight from the reference's formulas
    model, cfg = build({})
    T = model.n_timesteps
    s = torch.arange(T).float() / T
    tt = (torch.arange(T) + 1).float() / T
    g_s, g_t = model.gamma(s), model.gamma(tt)
    sigma2_ts, sigma_ts, alpha_ts, alpha_s = model.sigma_and_alpha_t_given_s(g_t, g_s)
    sigma_s, sigma_t = model.sigma(g_s), model.sigma(g_t)
    save["ep_c1"] = (alpha_ts * (sigma_s ** 2) / (sigma_t ** 2)).detach().numpy()
    save["ep_c2"] = (alpha_s * sigma2_ts / (sigma_t ** 2)).detach().numpy()
    np.savez_compressed(os.path.join(MG.GOLD, "endpoint_param.npz"), sizes=np.asarray(sizes), sample_sizes=np.asarray([4, 7]),
                        **save)
    print(os.path.getsize(os.path.join(MG.GOLD, "endpoint_param.npz")))


if __name__ == "__main__":
    main()
