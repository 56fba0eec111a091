"""The endpoint parameterisation (`endpoint_param_feat` / `endpoint_param_coord`, pharmacodiff.py:204-216, 413-418) and
`remove_com=False` (:123-125) on the GPU against the fixture written by the reference's own code with those flags
(`oracle/make_golden_endpoint.py` -> tests/golden/endpoint_param.npz) and against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PH = ["Aromatic", "HydrogenDonor", "HydrogenAcceptor", "PositiveIon", "NegativeIon", "Hydrophobic"]


def t(a):
    return torch.from_numpy(np.asarray(a))


def _model(sd, dyn_cfg, dropout=0.0, **flags):
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    cfg = dict(dyn_cfg, dropout=dropout)
    gcut = cfg.pop("graph_cutoffs")
    m = PharmacophoreDiff(6, 11, PH, n_timesteps=100, graph_config={"graph_cutoffs": gcut}, dynamics_config=cfg,
                          precision=1e-5, lr_scheduler_config={"base_lr": 1e-3, "weight_decay": 0.0}, **flags)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_posterior_step_endpoint_bit_exact(mode):
    """pf_posterior_step_ep against the same tensor arithmetic in torch (every product and sum rounded separately)."""
    from pharmacoforge_b200 import ops
    gen = torch.Generator().manual_seed(mode)
    sizes, atoms = [3, 8, 5, 16], [40, 7, 100, 33]
    fptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    pptr = torch.tensor(np.concatenate([[0], np.cumsum(atoms)]), dtype=torch.int32)
    nf, npr = int(fptr[-1]), int(pptr[-1])
    x, h = torch.randn(nf, 3, generator=gen) * 4, torch.randn(nf, 6, generator=gen)
    px, ph = torch.randn(nf, 3, generator=gen), torch.randn(nf, 6, generator=gen)
    nx, nh = torch.randn(nf, 3, generator=gen), torch.randn(nf, 6, generator=gen)
    prot = torch.randn(npr, 3, generator=gen) * 10
    a, v, q, c1, c2 = 0.9865, 0.0403, 0.1608, 0.9731, 0.0262
    f32 = lambda s: torch.tensor(s, dtype=torch.float32)
    mu_x = f32(c1) * x + f32(c2) * px if mode & 1 else x / f32(a) - f32(v) * px
    mu_h = f32(c1) * h + f32(c2) * ph if mode & 2 else h / f32(a) - f32(v) * ph
    zx, zh = mu_x + f32(q) * nx, mu_h + f32(q) * nh
    fb = torch.repeat_interleave(torch.arange(4), torch.tensor(sizes))
    pb = torch.repeat_interleave(torch.arange(4), torch.tensor(atoms))
    com = torch.zeros(4, 3).index_add_(0, fb, zx) / torch.tensor(sizes).float().view(-1, 1)
    gx, gh, gp = x.cuda(), h.cuda(), prot.cuda()
    ops.posterior_step_ep(gx, gh, px.cuda(), ph.cuda(), nx.cuda(), nh.cuda(), fptr.cuda(), gp, pptr.cuda(), a, v, q, c1, c2, mode)
    assert torch.equal(gh.cpu(), zh)
    assert torch.equal(gx.cpu(), zx - com[fb])
    assert torch.equal(gp.cpu(), prot - com[pb])


@pytest.mark.parametrize("tag,flags", [("ep", dict(endpoint_param_feat=True, endpoint_param_coord=True)),
                                       ("eph", dict(endpoint_param_feat=True))])
def test_endpoint_sampling_against_reference(golden, sd, dyn_cfg, tag, flags):
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    d = golden("endpoint_param.npz")
    sizes = [int(v) for v in d["sample_sizes"]]
    model = _model(sd, dyn_cfg, **flags).eval()
    noise = t(d[f"s_{tag}__noise"])
    for n in (1, 6):     # the first steps of the reference's own run (state in the COM-free frame, before the frame restore)
        g = GraphBatch.from_pockets([Pocket.from_numpy(*make_pocket(100, seed=5))], [sizes], "cuda:0")
        model.sample_given_receptor(g, noise=noise, n_steps=n, return_tensors=True)
        for got, ref in ((g.pharm_x, d[f"s_{tag}__traj_x"][n]), (g.pharm_h, d[f"s_{tag}__traj_h"][n])):
            ref = t(ref)
            err = float((got.cpu() - ref).abs().max())
            assert err <= 1e-3 * max(1.0, float(ref.abs().max())), (tag, n, err)
    if tag == "ep":      # the all-endpoint chain is contractive: compare the final sample of all 100 steps
        g = GraphBatch.from_pockets([Pocket.from_numpy(*make_pocket(100, seed=5))], [sizes], "cuda:0")
        x0, h0 = model.sample_given_receptor(g, noise=noise, return_tensors=True)
        assert float((x0.cpu() - t(d["s_ep__final_x"])).abs().max()) <= 2e-3
        assert float((h0.cpu() - t(d["s_ep__final_h"])).abs().max()) <= 2e-3


@pytest.mark.parametrize("tag,flags", [("ep", dict(endpoint_param_feat=True, endpoint_param_coord=True)),
                                       ("epx", dict(endpoint_param_coord=True)), ("nocom", dict(remove_com=False)),
                                       ("ep_nocom", dict(endpoint_param_feat=True, endpoint_param_coord=True, remove_com=False))])
def test_forward_losses_and_gradients_against_reference(golden, sd, dyn_cfg, tag, flags):
    """PharmacophoreDiff.forward with the flags set: the fused evaluation path and the differentiable training path against
    the reference's losses / metrics, and the gradient norm of every live parameter against the reference's backward."""
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    d = golden("endpoint_param.npz")
    sizes = [int(v) for v in d["sizes"]]
    model = _model(sd, dyn_cfg, **flags)
    inj = dict(t_int=t(d["t_int"]), eps={"x": t(d["eps_x"]), "h": t(d["eps_h"])})

    def batch():
        gb = GraphBatch.from_pockets([Pocket.from_numpy(*make_pocket(120, seed=5))], [sizes], "cuda:0")
        return gb.set_pharmacophores(t(d["x0"]), t(d["h0"]))

    def check(lo, me, phase):
        for k, v in {**lo, **me}.items():
            if "total" in k:
                continue
            ref = float(d[f"{tag}__{k.replace(phase, 'val').replace(' ', '_')}"])
            assert abs(float(v) - ref) <= 3e-4 * max(1.0, abs(ref)), (tag, phase, k, float(v), ref)
    model.eval()
    lo, me = model.validation_step(batch(), **inj)
    check(lo, me, "val")
    model.train()
    total, lo, me = model.training_step(batch(), **inj)
    check(lo, me, "train")
    total.backward()
    params = dict(model.named_parameters())
    for n, norm in zip(map(str, d[f"{tag}__grad_names"]), d[f"{tag}__grad_norms"]):
        gr = params[n].grad
        assert gr is not None, n
        assert abs(float(gr.double().norm()) - norm) <= 2e-3 * max(norm, 1e-6), (tag, n, float(gr.norm()), norm)


@pytest.mark.parametrize("tag", ["n4", "n1"])
def test_numeric_message_norm_against_reference(golden, sd, dyn_cfg, tag):
    """message_norm = a positive number (sum aggregation / norm, gvp.py:386-389, 512-517; the reference constructor's default is
    1): the fused denoiser -- per-layer features and eps -- and the differentiable training graph against a denoiser call of
    the reference's own code (oracle/make_golden_msgnorm.py)."""
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    d = golden("message_norm.npz")
    nv = float(d[f"{tag}__norm"])
    model = _model(sd, dict(dyn_cfg, message_norm=nv)).eval()
    sizes = [int(v) for v in d["sizes"]]
    g = GraphBatch.from_pockets([Pocket.from_numpy(*make_pocket(int(d["n_atoms"]), seed=int(d["pocket_seed"])))], [sizes], "cuda:0")
    dyn = model.dynamics
    st = dyn.bind(g)
    g.prot_x.copy_(t(d["prot_x"]))
    g.pharm_x.copy_(t(d["x_t"]))
    g.pharm_h.copy_(t(d["h_t"]))
    with torch.no_grad():
        eps_h, eps_x = dyn(g, t(d["t"]), None)

    errs = {}

    def close(a, ref, what, rtol=1e-4):
        ref = t(ref)
        err = float((a.cpu() - ref).abs().max())
        errs[what] = err / max(float(ref.abs().max()), 1e-6)
        assert err <= rtol * max(float(ref.abs().max()), 1e-6) + 1e-6, (tag, what, err)
    close(eps_h, d[f"{tag}__eps_h"], "eps_h")
    close(eps_x, d[f"{tag}__eps_x"], "eps_x")
    close(st.pharm_hh, d[f"{tag}__conv1_pharm_h"], "conv1 pharm h")
    close(st.prot_h, d[f"{tag}__conv1_prot_h"], "conv1 prot h")
    # vectors: ours are component-major [N, 3, 16], the reference's [N, 16, 3]
    close(st.prot_v.view(-1, 3, 16).transpose(1, 2), d[f"{tag}__conv1_prot_v"], "conv1 prot v")
    # differentiable graph (training mode, dropout 0): same eps
    from pharmacoforge_b200 import train_graph
    model.train()
    g.prot_x.copy_(t(d["prot_x"]))
    g.pharm_x.copy_(t(d["x_t"]))
    g.pharm_h.copy_(t(d["h_t"]))
    th, tx = train_graph.dynamics_forward(dyn, g, t(d["t"]), training=True)
    close(th.detach(), d[f"{tag}__eps_h"], "train eps_h", rtol=2e-4)
    close(tx.detach(), d[f"{tag}__eps_x"], "train eps_x", rtol=2e-4)
    (th.sum() + tx.sum()).backward()
    assert dyn.pharm_encoder[0].weight.grad is not None
