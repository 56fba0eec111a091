"""Configuration switches beyond configs/dev.yml on the GPU, each against a fixture written by the reference's own code with the
switch set (and against the oracle): the endpoint parameterisation (`endpoint_param_feat` / `endpoint_param_coord`,
pharmacodiff.py:204-216, 413-418) and `remove_com=False` (:123-125) -- `oracle/make_golden_endpoint.py` ->
tests/golden/endpoint_param.npz; a numeric `message_norm`, incl. 0 (gvp.py:375-389, 504-517) -- message_norm.npz /
message_norm0.npz; `pf_k = 0`, radius pf / fp edges (dynamics_gvp.py:210-216) -- pf_radius.npz."""
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


@pytest.mark.parametrize("tag", ["n4", "n1", "n0"])
def test_numeric_message_norm_against_reference(golden, sd, dyn_cfg, tag):
    """message_norm = a positive number (sum aggregation / norm, gvp.py:386-389, 512-517; the reference constructor's default is
    1): the fused denoiser -- per-layer features and eps -- and the differentiable training graph against a denoiser call of
    the reference's own code (oracle/make_golden_msgnorm.py)."""
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    d = golden("message_norm0.npz" if tag == "n0" else "message_norm.npz")   # n0: message_norm = 0, edges per node + 1
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


# ------------------------------------------------------------------------------------------------ pf_k == 0
def _canon(src, dst):
    s, d = src.cpu().numpy().astype(np.int64), dst.cpu().numpy().astype(np.int64)
    o = np.lexsort((s, d))
    return s[o], d[o]


def _radius_batch(pocket_specs, sizes, tile_rows=128, pf_max_nbrs=100):
    import pf_oracle as O
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    pk = [make_pocket(n, seed=s) for n, s in pocket_specs]
    g = GraphBatch.from_pockets([Pocket.from_numpy(p, h) for p, h in pk], sizes, "cuda:0", pf_k=0, tile_rows=tile_rows,
                                pf_max_nbrs=pf_max_nbrs)
    b = O.build_batch([(t(p), t(h)) for p, h in pk], sizes)
    return g, b


def _radius_graph(g, pf_r, ff_k=0):
    from pharmacoforge_b200 import ops
    ops.dyn_graph_radius(g.prot_x, g.prot_ptr, g.pharm_x, g.pharm_ptr, 9.0, g.ff_max_nbrs, ff_k, pf_r, g.pf_max_nbrs, g.tile_rows,
                         g.ff_start, g.ff_cnt, g.ff_col, g.pf_start, g.pf_sub_ptr, g.fp_base, g.pf_cnt, g.pf_col, g.pf_sub_start,
                         g.pf_sub_cnt, g.pf_sub_x, g.fp_seg_start, g.fp_seg_cnt, g.fp_col, g.status)
    g.check_status()


@pytest.mark.parametrize("pf_r,cap,tile_rows", [(8.0, 100, 128), (11.0, 100, 128), (30.0, 3, 128), (4.0, 100, 64)])
def test_pf_radius_graph_bit_exact(pf_r, cap, tile_rows):
    """pf_dyn_graph_radius (pf_k == 0, dynamics_gvp.py:210-216): ff / pf / fp edge lists bit-exact against the oracle on a ragged
    batch (a 1,500-atom pocket, a 3-atom pocket, a 1-node graph, a centre far from its pocket), the per-atom cap (cap = 3 with a
    radius that reaches everything: every atom keeps its three lowest-index nodes), and the sub-segment cut of the pf segments."""
    import pf_oracle as O
    g, b = _radius_batch([(400, 0), (1500, 3), (3, 9), (60, 7)], [[3, 8, 5], [6, 16], [4], [1, 12]], tile_rows, cap)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(g.n_pharm, 3, generator=gen) * 4.0
    x[1] = torch.tensor([60.0, 0.0, 0.0])
    com = torch.stack([b.prot_x[int(b.prot_ptr[i]):int(b.prot_ptr[i + 1])].mean(0) for i in range(b.n_graphs)])
    prot = b.prot_x - com[b.prot_b]
    g.pharm_x.copy_(x)
    g.prot_x.copy_(prot)
    b.pharm_x, b.prot_x = x, prot
    _radius_graph(g, pf_r)
    got = g.dynamic_edges()
    want = O.dynamic_edges(b, 9.0, 0, 0, pf_r)
    q, c = O.radius_bipartite_edges(b.prot_x, b.prot_ptr, b.pharm_x, b.pharm_ptr, pf_r, cap)
    want["pf"], want["fp"] = (q, c), (c, q)
    for et in ("ff", "pf", "fp"):
        s, d = _canon(*got[et])
        so, do = _canon(*want[et])
        assert np.array_equal(s, so) and np.array_equal(d, do), et
    # both lists come out destination-major with ascending sources, as radius() orders them
    assert np.array_equal(got["fp"][0].cpu().numpy(), c.numpy()) and np.array_equal(got["fp"][1].cpu().numpy(), q.numpy())
    n = g.n_pharm
    cnt, sub_ptr = g.pf_cnt[:n].cpu().numpy(), g.pf_sub_ptr.cpu().numpy()
    sub_cnt, sub_start = g.pf_sub_cnt.cpu().numpy(), g.pf_sub_start.cpu().numpy()
    pf_start = g.pf_start.cpu().numpy()
    assert sub_ptr[-1] == g.n_pf_sub and sub_cnt.max() <= tile_rows
    for i in range(n):
        sl = slice(sub_ptr[i], sub_ptr[i + 1])
        assert sub_cnt[sl].sum() == cnt[i]
        assert np.array_equal(sub_start[sl], pf_start[i] + tile_rows * np.arange(sub_ptr[i + 1] - sub_ptr[i]))
    if pf_r >= 8.0 and cap == 100:
        assert cnt.max() > tile_rows           # the case the sub-segments exist for
    if cap == 3:
        assert int(g.fp_seg_cnt[:g.n_prot].max()) == 3


@pytest.mark.parametrize("tile_rows", [128, 64])
def test_denoiser_radius_pf_edges_against_reference(golden, sd, dyn_cfg, tile_rows):
    """dynamics_config pf_k = 0: the fused denoiser (tcgen05 and FFMA kernels) and the differentiable training graph against a
    denoiser call of the reference's own code with that switch (oracle/make_golden_pfradius.py: pf cutoff 11 A on a 400-atom
    pocket, pharmacophore in-degrees up to ~400 = four tiles), edge lists bit-exact, per-layer features and eps."""
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    d = golden("pf_radius.npz")
    cuts = dict(dyn_cfg["graph_cutoffs"], pf=float(d["pf_cutoff"]), fp=float(d["pf_cutoff"]))
    model = _model(sd, dict(dyn_cfg, pf_k=0, graph_cutoffs=cuts)).eval()
    sizes = [int(v) for v in d["sizes"]]
    g = GraphBatch.from_pockets([Pocket.from_numpy(*make_pocket(int(d["n_atoms"]), seed=int(d["pocket_seed"])))], [sizes], "cuda:0",
                                pf_k=0, tile_rows=tile_rows)
    dyn = model.dynamics
    st = dyn.bind(g)

    def load():
        g.prot_x.copy_(t(d["prot_x"]))
        g.pharm_x.copy_(t(d["x_t"]))
        g.pharm_h.copy_(t(d["h_t"]))
    load()
    with torch.no_grad():
        eps_h, eps_x = dyn(g, t(d["t"]), None)
    got = g.dynamic_edges()
    for et in ("ff", "pf", "fp"):
        s_, d_ = _canon(*got[et])
        assert np.array_equal(s_, d[f"e_{et}_src"]) and np.array_equal(d_, d[f"e_{et}_dst"]), et
    assert int(g.pf_cnt.max()) > 128

    def close(a, ref, what, rtol=1e-4):
        ref = t(ref)
        err = float((a.cpu() - ref).abs().max())
        print(tile_rows, what, "max abs err", err, "relative to max", err / max(float(ref.abs().max()), 1e-6))
        assert err <= rtol * max(float(ref.abs().max()), 1e-6) + 1e-6, (what, err)
    close(eps_h, d["eps_h"], "eps_h")
    close(eps_x, d["eps_x"], "eps_x")
    close(st.pharm_hh, d["conv1_pharm_h"], "conv1 pharm h")
    close(st.prot_h, d["conv1_prot_h"], "conv1 prot h")
    if tile_rows == 128:
        # every layer-0 switch gives the same eps (the table / seeded first layer run on the sub-segment list too)
        for attr in ("layer0_table", "layer0_seed"):
            setattr(dyn, attr, False)
            load()
            with torch.no_grad():
                eh2, ex2 = dyn(g, t(d["t"]), None)
            close(eh2, d["eps_h"], f"eps_h ({attr} off)")
            close(ex2, d["eps_x"], f"eps_x ({attr} off)")
        dyn.layer0_table = dyn.layer0_seed = True
        # differentiable graph (training mode, dropout 0): same eps, gradients reach the encoders
        from pharmacoforge_b200 import train_graph
        model.train()
        load()
        th, tx = train_graph.dynamics_forward(dyn, g, t(d["t"]), training=True)
        close(th.detach(), d["eps_h"], "train eps_h", rtol=2e-4)
        close(tx.detach(), d["eps_x"], "train eps_x", rtol=2e-4)
        (th.sum() + tx.sum()).backward()
        assert dyn.prot_encoder[0].weight.grad is not None and torch.isfinite(dyn.prot_encoder[0].weight.grad).all()


def test_radius_pf_edges_numeric_norm_and_sampling(sd, dyn_cfg):
    """pf_k = 0 together with a numeric message_norm (10, and 0) against the oracle, and a short reverse diffusion (8 steps) whose result is
    bit-identical between one batch and batches of one graph (the sub-segment combine keeps the batch-composition invariance)."""
    import pf_oracle as O
    cuts = dict(dyn_cfg["graph_cutoffs"], pf=8.0, fp=8.0)
    for mn in (10.0, 0):      # 0: edges per node of the graph + 1, true per-graph counts in this mode (pf_degree_norms)
        cfg = dict(dyn_cfg, pf_k=0, message_norm=mn, graph_cutoffs=cuts)
        model = _model(sd, cfg).eval()
        g, b = _radius_batch([(400, 1), (150, 4)], [[4, 7], [5]])
        gen = torch.Generator().manual_seed(17)
        x, h = torch.randn(g.n_pharm, 3, generator=gen) * 3.0, torch.randn(g.n_pharm, 6, generator=gen)
        com = torch.stack([b.prot_x[int(b.prot_ptr[i]):int(b.prot_ptr[i + 1])].mean(0) for i in range(b.n_graphs)])
        prot = b.prot_x - com[b.prot_b]
        model.dynamics.bind(g)
        g.pharm_x.copy_(x)
        g.pharm_h.copy_(h)
        g.prot_x.copy_(prot)
        b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
        tt = torch.tensor([0.3, 0.3, 0.75])
        wh, wx = O.denoiser(sd, b, tt, cfg)
        with torch.no_grad():
            gh, gx = model.dynamics(g, tt, None)
        for a_, w_, what in ((gh, wh, "eps_h"), (gx, wx, "eps_x")):
            err = float((a_.cpu() - w_).abs().max())
            assert err <= 1e-4 * max(float(w_.abs().max()), 1e-6) + 1e-6, (mn, what, err)
    # sampling: the second pocket alone gives bit-identical results (a graph never sees its batch neighbours)
    from pharmacoforge_b200.batch import Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    model = _model(sd, dict(dyn_cfg, pf_k=0, graph_cutoffs=cuts)).eval()
    pockets = [Pocket.from_numpy(*make_pocket(n, seed=s_)) for n, s_ in ((400, 1), (150, 4))]
    noise = torch.randn(9, 16, 9, generator=torch.Generator().manual_seed(3))
    x1, h1 = model.sample_given_receptor(model.make_batch(pockets, [[4, 7], [5]], "cuda:0"), noise=noise, n_steps=8,
                                         return_tensors=True)
    x2, h2 = model.sample_given_receptor(model.make_batch(pockets[1:], [[5]], "cuda:0"), noise=noise[:, 11:].contiguous(),
                                         n_steps=8, return_tensors=True)
    assert torch.isfinite(x1).all() and torch.isfinite(h1).all()
    assert torch.equal(x2, x1[11:]) and torch.equal(h2, h1[11:])
    # and the Philox / CUDA-graph throughput path in this mode: the captured loop replays the eager one bit for bit
    out = model.sample_given_receptor(model.make_batch(pockets, [[4, 7], [5]], "cuda:0"))
    assert len(out) == 3 and all(torch.isfinite(o.ph_coords).all() and torch.isfinite(o.ph_feats).all() for o in out)
    gb = model.make_batch(pockets, [[4, 7], [5]], "cuda:0")

    def run(graph):
        model.use_cuda_graph = graph
        torch.manual_seed(5)
        gb.prot_x.copy_(gb.prot_x0)
        xx, hh = model.sample_given_receptor(gb, n_steps=10, return_tensors=True)
        return xx.clone(), hh.clone()
    try:
        e1 = run(False)
        g1, g2 = run(True), run(True)      # captures, then replays
        assert len(model.dynamics.bind(gb).graphs) == 1
        for r in (g1, g2):
            assert torch.equal(r[0], e1[0]) and torch.equal(r[1], e1[1])
    finally:
        model.use_cuda_graph = False
