"""Pin oracle/pf_oracle.py (the CPU restatement that travels to the GPU box) against fixtures produced
by the reference's own code (oracle/make_golden.py).  CPU only."""
import numpy as np
import torch

import pf_oracle as O
from pharmacoforge_b200.synthetic import make_pocket


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_known_answer_constants(golden):
    c = golden("constants.npz")
    g = O.gamma_table(100, 1e-5)
    assert torch.equal(g, t(c["gamma"]))
    # SURVEY.md App. A.6
    ref = [-11.51291561, -8.46825981, -0.25130934, 7.80874634, 11.47407913]
    assert np.allclose(g[[0, 1, 50, 99, 100]].numpy(), ref, rtol=0, atol=1e-6)
    T = 100
    s = torch.arange(T).float() / T
    tt = (torch.arange(T) + 1).float() / T
    a, v, q = O.posterior_coefficients(O.gamma_at(g, s, T), O.gamma_at(g, tt, T))
    assert torch.equal(a, t(c["alpha_ts"])) and torch.equal(v, t(c["var_terms"])) and torch.equal(q, t(c["sigma_q"]))
    assert np.allclose([a[99], v[99], q[99]], [0.16001807, 6.08930731, 0.98691881], atol=1e-6)
    assert np.allclose([a[0], v[0], q[0]], [0.99989998, 0.01380232, 0.00308608], atol=1e-6)
    assert torch.equal(O.rbf(torch.tensor([3.0])), t(c["rbf3"]))
    assert torch.equal(O.rbf(torch.linspace(0, 20, 41)), t(c["rbf_grid"]))
    assert torch.equal(O.norm_no_nan(torch.zeros(1, 3)), t(c["norm0"]))
    assert abs(float(c["norm0"][0]) - 1e-4) < 1e-9


def test_pp_radius_graph(golden):
    gpp = golden("pp_graph_n400_seed0.npz")
    pos, _ = make_pocket(400, seed=0)
    src, dst = O.radius_edges(t(pos), torch.tensor([0, 400]), 3.5, 100)
    assert np.array_equal(src.numpy(), gpp["src"]) and np.array_equal(dst.numpy(), gpp["dst"])
    deg = np.bincount(gpp["dst"], minlength=400)
    assert 2900 < src.numel() < 3100 and deg.max() <= 18


def _denoiser_batch(d):
    sizes = [int(v) for v in d["sizes"]]
    pos, onehot = make_pocket(int(d["n_atoms"]), seed=int(d["pocket_seed"]))
    b = O.build_batch([(t(pos), t(onehot))], [sizes])
    b.prot_x = t(d["prot_x"]).clone()
    b.pharm_x = t(d["x_t"]).clone()
    b.pharm_h = t(d["h_t"]).clone()
    return b


def _canon(src, dst):
    o = np.lexsort((src.numpy(), dst.numpy()))
    return src.numpy()[o], dst.numpy()[o]


def test_denoiser_call(golden, sd, dyn_cfg):
    d = golden("denoiser_call.npz")
    b = _denoiser_batch(d)
    trace = {}
    eps_h, eps_x = O.denoiser(sd, b, t(d["t"]), dyn_cfg, trace=trace)
    for et in ("ff", "pf", "fp", "pp"):
        s_, d_ = _canon(*trace["edges"][et])
        assert np.array_equal(s_, d[f"e_{et}_src"]) and np.array_equal(d_, d[f"e_{et}_dst"]), et
    tol = dict(rtol=1e-5, atol=1e-6)
    assert np.allclose(trace["enc"]["pharm"].numpy(), d["enc_pharm"], **tol)
    assert np.allclose(trace["enc"]["prot"].numpy(), d["enc_prot"], **tol)
    for li in (0, 1):
        for nt in ("pharm", "prot"):
            h, v = trace[f"conv{li}"][nt]
            assert np.allclose(h.numpy(), d[f"conv{li}_{nt}_h"], **tol), (li, nt)
            assert np.allclose(v.numpy(), d[f"conv{li}_{nt}_v"], **tol), (li, nt)
    assert np.allclose(eps_h.numpy(), d["eps_h"], **tol)
    assert np.allclose(eps_x.numpy(), d["eps_x"], **tol)


def test_full_reverse_diffusion(golden, sd, dyn_cfg):
    d = golden("sample_traj.npz")
    sizes = [int(v) for v in d["sizes"]]
    pos, onehot = make_pocket(int(d["n_atoms"]), seed=int(d["pocket_seed"]))
    b = O.build_batch([(t(pos), t(onehot))], [sizes])
    rec = []
    x0, h0, types, prot = O.sample(sd, b, t(d["noise"]), 100, sd["gamma.gamma"], dyn_cfg, record=rec)
    assert len(rec) == 101
    # same ops in the same order on the same machine: the trajectory must track the reference closely
    # for all 100 steps (no neighbour-list flip occurs on this seed)
    tx = torch.stack([r[0] for r in rec]).numpy()
    th = torch.stack([r[1] for r in rec]).numpy()
    assert np.allclose(tx, d["traj_x"], rtol=1e-4, atol=1e-4)
    assert np.allclose(th, d["traj_h"], rtol=1e-4, atol=1e-4)
    tp = torch.stack([r[2][[0, 100]] for r in rec]).numpy()
    assert np.allclose(tp, d["traj_prot0"], rtol=1e-4, atol=1e-4)
    assert np.allclose(x0.numpy(), d["final_x"], rtol=1e-4, atol=1e-4)
    assert np.allclose(h0.numpy(), d["final_h"], rtol=1e-4, atol=1e-4)
    assert np.array_equal(types.numpy(), d["final_type"])
    assert np.allclose(prot.numpy(), d["final_prot"], rtol=1e-4, atol=1e-3)


def _sample_multi_inputs(d):
    pockets = [tuple(t(a) for a in make_pocket(int(n), seed=int(s))) for n, s in d["pockets"]]
    flat, n_pharms, i = d["n_pharms_flat"].tolist(), [], 0
    for k in d["n_pharms_per_pocket"].tolist():
        n_pharms.append(flat[i:i + k])
        i += k
    return pockets, n_pharms


def test_sample_multi_and_trajectory_frames(golden, sd, dyn_cfg):
    """The reference's own `PharmacophoreDiff.sample` (pharmacodiff.py:516-578) with max_batch_size chunking, an explicit
    per-pocket init_pharm_com and visualize_trajectory=True: the oracle's restatement reproduces the final samples and
    the VALUES of all 101 trajectory frames (get_pos_feat_for_visual, :360-378)."""
    d = golden("sample_multi.npz")
    pockets, n_pharms = _sample_multi_inputs(d)
    out = O.sample_multi(sd, pockets, n_pharms, t(d["noise"]), 100, sd["gamma.gamma"], dyn_cfg,
                         max_batch_size=int(d["max_batch_size"]), init_pharm_com=t(d["init_pharm_com"]), frames=True)
    assert [len(o) for o in out] == d["n_pharms_per_pocket"].tolist()
    ph = [p for o in out for p in o]
    assert [p["x"].shape[0] for p in ph] == d["n_pharms_flat"].tolist()
    assert np.allclose(torch.cat([p["x"] for p in ph]).numpy(), d["final_x"], rtol=1e-4, atol=1e-4)
    assert np.allclose(torch.cat([p["h"] for p in ph]).numpy(), d["final_h"], rtol=1e-4, atol=1e-4)
    assert np.array_equal(torch.cat([p["type"] for p in ph]).numpy(), d["final_type"])
    assert np.allclose(torch.cat([p["pos_frames"] for p in ph], 1).numpy(), d["pos_frames"], rtol=1e-4, atol=1e-4)
    assert np.allclose(torch.cat([p["feat_frames"] for p in ph], 1).numpy(), d["feat_frames"], rtol=1e-4, atol=1e-4)


def test_forward_loss_matches_reference_forward(golden, sd, dyn_cfg):
    """oracle.forward_loss against PharmacophoreDiff.forward of the reference run with injected (t, eps)
    (oracle/make_golden_loss.py -> tests/golden/forward_loss.npz)."""
    g = golden("forward_loss.npz")
    pos, onehot = make_pocket(int(g["n_atoms"]), seed=int(g["pocket_seed"]))
    b = O.build_batch([(t(pos), t(onehot))], [list(map(int, g["sizes"]))])
    losses, metrics = O.forward_loss(sd, b, t(g["x0"]), t(g["h0"]), t(g["t_int"]), t(g["eps_x"]), t(g["eps_h"]), 100,
                                     sd["gamma.gamma"], dyn_cfg, phase="val")
    for k, v in {**losses, **metrics}.items():
        ref = float(g[k.replace(" ", "_")])
        assert abs(float(v) - ref) <= 1e-5 * max(1.0, abs(ref)), (k, float(v), ref)


def test_oracle_gradients_match_reference_backward(golden, sd, dyn_cfg):
    """Autograd through oracle.forward_loss against the reference's own backward of the same loss (total = pos + feat,
    pharmacodiff.py:265-297): gradient norm and sum of every parameter, the full gradient of five of them, and the set
    of parameters that get no gradient at all (protein side of the last conv layer)."""
    g = golden("forward_loss.npz")
    pos, onehot = make_pocket(int(g["n_atoms"]), seed=int(g["pocket_seed"]))
    b = O.build_batch([(t(pos), t(onehot))], [list(map(int, g["sizes"]))])
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != "gamma.gamma" and v.numel() else v)
           for k, v in sd.items()}
    losses, _ = O.forward_loss(sdg, b, t(g["x0"]), t(g["h0"]), t(g["t_int"]), t(g["eps_x"]), t(g["eps_h"]), 100,
                               sd["gamma.gamma"], dyn_cfg, phase="val")
    torch.stack(list(losses.values())).sum().backward()
    names = [str(n) for n in g["grad_names"]]
    for n, norm, total in zip(names, g["grad_norms"], g["grad_sums"]):
        gr = sdg[n].grad
        assert gr is not None, n
        assert abs(float(gr.double().norm()) - norm) <= 2e-4 * max(norm, 1e-6), (n, float(gr.norm()), norm)
        assert abs(float(gr.double().sum()) - total) <= 2e-4 * max(norm, 1e-6) * gr.numel() ** 0.5, n
    for n in map(str, g["dead_params"]):
        assert sdg[n].grad is None or float(sdg[n].grad.abs().max()) == 0.0, n
    for k in g:
        if k.startswith("grad__"):
            ref = t(g[k])
            got = sdg[k[6:]].grad
            assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max()), k


# ---- the endpoint parameterisation and remove_com=False (outside configs/dev.yml; README.md names an endpoint_param.yaml)
import pytest  # noqa: E402


@pytest.mark.parametrize("tag,flags", [("ep", dict(endpoint_feat=True, endpoint_coord=True)),
                                       ("epx", dict(endpoint_coord=True)), ("nocom", dict(remove_com=False)),
                                       ("ep_nocom", dict(endpoint_feat=True, endpoint_coord=True, remove_com=False))])
def test_forward_loss_endpoint_and_nocom_against_reference(golden, sd, dyn_cfg, tag, flags):
    g = golden("endpoint_param.npz")
    pos, onehot = make_pocket(120, seed=5)
    b = O.build_batch([(t(pos), t(onehot))], [list(map(int, g["sizes"]))])
    lo, me = O.forward_loss(sd, b, t(g["x0"]), t(g["h0"]), t(g["t_int"]), t(g["eps_x"]), t(g["eps_h"]), 100,
                            sd["gamma.gamma"], dyn_cfg, phase="val", **flags)
    for k, v in {**lo, **me}.items():
        ref = float(g[f"{tag}__{k.replace(' ', '_')}"])
        assert abs(float(v) - ref) <= 2e-5 * max(1.0, abs(ref)), (tag, k, float(v), ref)


@pytest.mark.parametrize("tag,flags", [("ep", dict(endpoint_feat=True, endpoint_coord=True)), ("eph", dict(endpoint_feat=True))])
def test_endpoint_reverse_diffusion_against_reference(golden, sd, dyn_cfg, tag, flags):
    """The endpoint posterior (pharmacodiff.py:413-418): coefficient tables bit-identical to the reference's formulas, the
    first reverse steps of the reference's own run reproduced from its injected noise, and -- for the all-endpoint model,
    whose chain is contractive -- the final sample."""
    g = golden("endpoint_param.npz")
    gam = sd["gamma.gamma"]
    T = 100
    s = torch.arange(T).float() / T
    tt = (torch.arange(T) + 1).float() / T
    c1, c2 = O.endpoint_coefficients(O.gamma_at(gam, s, T), O.gamma_at(gam, tt, T))
    assert np.array_equal(c1.numpy(), g["ep_c1"]) and np.array_equal(c2.numpy(), g["ep_c2"])
    pos, onehot = make_pocket(100, seed=5)
    sizes = list(map(int, g["sample_sizes"]))
    noise, tx, th = (t(g[f"s_{tag}__{k}"]) for k in ("noise", "traj_x", "traj_h"))
    b = O.build_batch([(t(pos), t(onehot))], [sizes])
    rec = []
    n = 100 if tag == "ep" else 6
    x0, h0, _, _ = O.sample(sd, b, noise, T, gam, dyn_cfg, record=rec, steps=n, endpoint_feat=flags.get("endpoint_feat", False),
                            endpoint_coord=flags.get("endpoint_coord", False))
    for i in range(0, n + 1, max(1, n // 6)):
        sx, sh = max(1.0, float(tx[i].abs().max())), max(1.0, float(th[i].abs().max()))
        assert float((rec[i][0] - tx[i]).abs().max()) <= 1e-3 * sx, (tag, i, float((rec[i][0] - tx[i]).abs().max()))
        assert float((rec[i][1] - th[i]).abs().max()) <= 1e-3 * sh, (tag, i)
    if tag == "ep":
        assert float((x0 - t(g["s_ep__final_x"])).abs().max()) <= 2e-3
        assert float((h0 - t(g["s_ep__final_h"])).abs().max()) <= 2e-3


@pytest.mark.parametrize("tag", ["n4", "n1", "n0"])
def test_denoiser_numeric_message_norm_against_reference(golden, sd, dyn_cfg, tag):
    """message_norm = a positive number: sum aggregation divided by it (gvp.py:386-389, 512-517) -- the constructor default
    of the reference's dynamics is 1; 0: divided by edges per node of the graph + 1 (:504-507), with the per-graph edge counts
    as dynamics_gvp.py:219-221 records them.  Fixtures from the reference's own code (oracle/make_golden_msgnorm.py)."""
    d = golden("message_norm0.npz" if tag == "n0" else "message_norm.npz")
    b = _denoiser_batch(d)
    cfg = dict(dyn_cfg, message_norm=float(d[f"{tag}__norm"]))
    trace = {}
    eps_h, eps_x = O.denoiser(sd, b, t(d["t"]), cfg, trace=trace)
    for li in range(2):
        for nt in ("pharm", "prot"):
            h, v = trace[f"conv{li}"][nt]
            ref_h, ref_v = t(d[f"{tag}__conv{li}_{nt}_h"]), t(d[f"{tag}__conv{li}_{nt}_v"])
            assert float((h - ref_h).abs().max()) <= 2e-5 * max(1.0, float(ref_h.abs().max())), (tag, li, nt)
            assert float((v - ref_v).abs().max()) <= 2e-5 * max(1.0, float(ref_v.abs().max())), (tag, li, nt)
    assert float((eps_h - t(d[f"{tag}__eps_h"])).abs().max()) <= 2e-5
    assert float((eps_x - t(d[f"{tag}__eps_x"])).abs().max()) <= 2e-5


def test_denoiser_radius_pf_edges_against_reference(golden, sd, dyn_cfg):
    """pf_k = 0: pf / fp edges from radius(pharm, prot, r_pf, max 100 per protein atom) (dynamics_gvp.py:210-216): edge sets
    bit-exact, features and eps against a denoiser call of the reference's own code (oracle/make_golden_pfradius.py)."""
    d = golden("pf_radius.npz")
    b = _denoiser_batch(d)
    cuts = dict(dyn_cfg["graph_cutoffs"], pf=float(d["pf_cutoff"]), fp=float(d["pf_cutoff"]))
    cfg = dict(dyn_cfg, pf_k=0, graph_cutoffs=cuts)
    trace = {}
    eps_h, eps_x = O.denoiser(sd, b, t(d["t"]), cfg, trace=trace)
    for et in ("ff", "pf", "fp"):
        s_, d_ = _canon(*trace["edges"][et])
        assert np.array_equal(s_, d[f"e_{et}_src"]) and np.array_equal(d_, d[f"e_{et}_dst"]), et
    assert np.bincount(d["e_pf_dst"]).max() > 128          # the fixture exercises destinations wider than one tile
    for nt in ("pharm", "prot"):
        h = trace["conv1"][nt][0]
        ref = t(d[f"conv1_{nt}_h"])
        assert float((h - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max())), nt
    assert float((eps_h - t(d["eps_h"])).abs().max()) <= 2e-5
    assert float((eps_x - t(d["eps_x"])).abs().max()) <= 2e-5
