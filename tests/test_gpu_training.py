"""Training path on the GPU: every differentiable custom op (forward + backward kernels) against a plain PyTorch fp32
reference of the same op, and the end-to-end gradients of PharmacophoreDiff.training_step against the reference's own
backward (golden fixture) and the oracle's autograd."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def t(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=2e-4, what=""):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(float(b.abs().max()), 1e-6)
    err = float((a - b).abs().max())
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


def check_op(fn, ref, inputs, what):
    """fn / ref take the same tensors; compares outputs and the gradients of a random cotangent w.r.t. every
    floating-point input that requires grad."""
    a = [x.clone().requires_grad_(x.is_floating_point()) if torch.is_tensor(x) else x for x in inputs]
    b = [x.clone().requires_grad_(x.is_floating_point()) if torch.is_tensor(x) else x for x in inputs]
    ya, yb = fn(*a), ref(*b)
    close(ya, yb, what=what + " fwd")
    cot = torch.randn_like(yb)
    ya.backward(cot)
    yb.backward(cot)
    for i, (p, q) in enumerate(zip(a, b)):
        if torch.is_tensor(q) and q.is_floating_point():
            assert p.grad is not None, (what, i)
            close(p.grad, q.grad, what=f"{what} grad[{i}]")


@pytest.fixture(scope="module")
def T():
    from pharmacoforge_b200 import train_ops
    return train_ops


def rnd(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).cuda()


@pytest.mark.parametrize("M,K,N,bias", [(1, 7, 128, True), (200, 161, 128, True), (1000, 144, 128, True),
                                        (999, 17, 17, False), (3001, 128, 16, True), (70000, 12, 128, True)])
def test_linear(T, M, K, N, bias):
    x, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    b = rnd(N, seed=3) if bias else None
    check_op(lambda x, w, *b: T.linear(x, w, b[0] if b else None), lambda x, w, *b: torch.nn.functional.linear(x, w, *b),
             [x, w] + ([b] if bias else []), f"linear {M}x{K}x{N}")


def test_silu_gate_vecnorm(T):
    x = rnd(777, 128, seed=4, scale=3.0)
    check_op(T.silu, torch.nn.functional.silu, [x], "silu")
    g, vu = rnd(333, 16, seed=5, scale=2.0), rnd(333, 3, 16, seed=6)
    check_op(lambda g, v: T.gate(g, v, True), lambda g, v: torch.sigmoid(g)[:, None, :] * v, [g, vu], "gate sigmoid")
    g1, vu1 = rnd(50, 1, seed=7), rnd(50, 3, 1, seed=8)
    check_op(lambda g, v: T.gate(g, v, False), lambda g, v: g[:, None, :] * v, [g1, vu1], "gate identity")
    vh = rnd(500, 3, 17, seed=9)
    vh[::7] = 0.0                      # rows below the clamp: zero gradient, value sqrt(1e-8)
    ref = lambda v: torch.sqrt(torch.clamp(v.square().sum(dim=1), min=1e-8))
    check_op(T.vecnorm, ref, [vh], "vecnorm")


def test_layernorms(T):
    x, w, b = rnd(1234, 128, seed=10, scale=2.0), rnd(128, seed=11), rnd(128, seed=12)
    check_op(T.layernorm, lambda x, w, b: torch.nn.functional.layer_norm(x, (128,), w, b, 1e-5), [x, w, b], "layernorm")
    v = rnd(600, 3, 16, seed=13)
    v[::5] *= 1e-6

    def ref(v):
        vn = torch.clamp(v.square().sum(dim=1, keepdim=True), min=1e-8)          # [M,1,U]
        vn = torch.sqrt(vn.mean(dim=2, keepdim=True) + 1e-5) + 1e-5
        return v / vn
    check_op(T.vecln, ref, [v], "vector layernorm")


def test_gather_segmean_geom(T):
    gen = torch.Generator().manual_seed(14)
    x = rnd(100, 3, 16, seed=15)
    idx = torch.randint(0, 100, (999,), generator=gen).int().cuda()
    check_op(lambda x: T.gather(x, idx), lambda x: x[idx.long()], [x], "gather")
    cnt = torch.randint(0, 9, (60,), generator=gen)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(cnt, 0)]).int().cuda()
    E = int(cnt.sum())
    seg_dst = torch.randperm(80, generator=gen)[:60].int().cuda()
    msg = rnd(E, 128, seed=16)
    seg_of = torch.repeat_interleave(torch.arange(60), cnt).cuda()

    def ref(m, dst):
        out = torch.zeros(80, 128, device="cuda").index_add(0, dst[seg_of], m)
        c = torch.zeros(80, device="cuda").index_add(0, dst[seg_of], torch.ones(E, device="cuda")).clamp(min=1)
        return out / c[:, None]
    check_op(lambda m: T.segmean(m, ptr, seg_dst, 80), lambda m: ref(m, seg_dst.long()), [msg], "segmean seg_dst")
    ar = torch.arange(60).cuda()
    check_op(lambda m: T.segmean(m, ptr, None, 60), lambda m: ref(m, ar)[:60], [msg], "segmean implicit")
    sx, dx = rnd(40, 3, seed=17, scale=4.0), rnd(30, 3, seed=18, scale=4.0)
    s = torch.randint(0, 40, (500,), generator=gen).int().cuda()
    d = torch.randint(0, 30, (500,), generator=gen).int().cuda()
    xd, rbf = T.edge_geom(sx, dx, s, d)
    diff = sx[s.long()] - dx[d.long()]
    dist = torch.sqrt(torch.clamp(diff.square().sum(1, keepdim=True), min=1e-8)) + 1e-8
    close(xd, diff / dist, what="x_diff")
    mu = torch.linspace(0, 15, 16, device="cuda")
    close(rbf, torch.exp(-((dist - mu) / (15 / 16)) ** 2), what="rbf")


@pytest.mark.parametrize("vi,vo,n,no,act", [(17, 16, 144, 128, True), (16, 16, 128, 128, True), (16, 1, 128, 64, False)])
def test_fused_gvp_op_matches_the_unfused_composition(T, vi, vo, n, no, act):
    """T.gvp (one host call per GVP, forward and backward) against the same GVP composed of the single-kernel ops."""
    from pharmacoforge_b200 import train_graph
    from pharmacoforge_b200.dynamics import GVP
    torch.manual_seed(vi * 100 + no)
    m = GVP(vi, vo, n, no, vectors_activation=None if act else torch.nn.Identity()).cuda()
    M = 777
    feats, vec = rnd(M, n, seed=21), rnd(M, 3, vi, seed=22)
    outs = []
    for fn in (train_graph.gvp_forward, train_graph.gvp_forward_unfused):
        m.zero_grad()
        a, b = feats.clone().requires_grad_(True), vec.clone().requires_grad_(True)
        f, v = fn(m, a, b)
        (f * rnd(M, no, seed=23)).sum().add((v * rnd(M, 3, vo, seed=24)).sum()).backward()
        outs.append([f, v, a.grad, b.grad] + [p.grad.clone() for p in m.parameters()])
    for i, (x, y) in enumerate(zip(*outs)):
        close(x, y, rtol=5e-4, what=f"gvp output/grad {i}")


def _model(sd, dyn_cfg, dropout):
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    cfg = dict(dyn_cfg, dropout=dropout)
    gcut = cfg.pop("graph_cutoffs")
    m = PharmacophoreDiff(6, 11, ["a", "b", "c", "d", "e", "f"], n_timesteps=100, graph_config={"graph_cutoffs": gcut},
                          dynamics_config=cfg, precision=1e-5, lr_scheduler_config={"base_lr": 1e-3, "weight_decay": 0.0})
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def test_training_step_gradients_match_reference_backward(golden, sd, dyn_cfg):
    """training_step (dropout 0, injected t / eps) -> loss.backward(): losses against the reference's forward, and the
    gradient of every parameter against the reference's own backward (norms, sums, five full tensors) and the oracle."""
    import pf_oracle as O
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    g = golden("forward_loss.npz")
    sizes = list(map(int, g["sizes"]))
    pos, onehot = make_pocket(int(g["n_atoms"]), seed=int(g["pocket_seed"]))
    model = _model(sd, dyn_cfg, dropout=0.0).train()
    gb = GraphBatch.from_pockets([Pocket.from_numpy(pos, onehot)], [sizes], "cuda:0")
    gb.set_pharmacophores(t(g["x0"]), t(g["h0"]))
    total, losses, metrics = model.training_step(gb, t_int=t(g["t_int"]), eps={"x": t(g["eps_x"]), "h": t(g["eps_h"])})
    assert abs(float(losses["train pos loss"]) - float(g["val_pos_loss"])) < 2e-4
    assert abs(float(losses["train feat loss"]) - float(g["val_feat_loss"])) < 2e-4
    assert abs(float(metrics["train accuracy"]) - float(g["val_accuracy"])) < 1e-6
    total.backward()
    params = dict(model.named_parameters())
    for n, norm, tot in zip(map(str, g["grad_names"]), g["grad_norms"], g["grad_sums"]):
        gr = params[n].grad
        assert gr is not None, n
        assert abs(float(gr.double().norm()) - norm) <= 1e-3 * max(norm, 1e-6), (n, float(gr.norm()), norm)
    for n in map(str, g["dead_params"]):
        assert params[n].grad is None, n          # protein side of the last layer: unused, exactly like the reference
    for k in g:
        if k.startswith("grad__"):
            close(params[k[6:]].grad, t(g[k]), rtol=1e-3, what=k)
    # and against the oracle's autograd on every parameter, elementwise
    b = O.build_batch([(t(pos), t(onehot))], [sizes])
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != "gamma.gamma" and v.numel() else v)
           for k, v in sd.items()}
    lo, _ = O.forward_loss(sdg, b, t(g["x0"]), t(g["h0"]), t(g["t_int"]), t(g["eps_x"]), t(g["eps_h"]), 100,
                           sd["gamma.gamma"], dyn_cfg, phase="train")
    torch.stack(list(lo.values())).sum().backward()
    for n in map(str, g["grad_names"]):
        close(params[n].grad, sdg[n].grad, rtol=1e-3, what=n)


def test_adam_steps_reduce_the_loss_and_eval_path_follows(sd, dyn_cfg):
    """A few optimisation steps with dropout 0.1 (training-mode masks) on one batch lower its loss, and the fused
    evaluation kernels, re-packed from the updated parameters, agree with the differentiable path."""
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    torch.manual_seed(0)
    pockets = [Pocket.from_numpy(*make_pocket(150, seed=40 + i)) for i in range(3)]
    sizes = [[4], [6], [5]]
    gen = torch.Generator().manual_seed(3)
    nf = sum(s[0] for s in sizes)
    h0 = torch.nn.functional.one_hot(torch.randint(0, 6, (nf,), generator=gen), 6).float()
    model = _model(sd, dyn_cfg, dropout=0.1).train()
    opt = model.configure_optimizers()["optimizer"]
    t_int = torch.tensor([20, 50, 80])
    eps = {"x": torch.randn(nf, 3, generator=gen), "h": torch.randn(nf, 6, generator=gen)}

    def batch():
        gb = GraphBatch.from_pockets(pockets, sizes, "cuda:0")
        x0 = torch.cat([p.prot_x.mean(0, keepdim=True) + torch.randn(s[0], 3, generator=torch.Generator().manual_seed(9))
                        for p, s in zip(pockets, sizes)])
        return gb.set_pharmacophores(x0, h0)

    def eval_loss():
        model.eval()
        lo, _ = model.validation_step(batch(), t_int=t_int, eps=eps)
        model.train()
        return float(lo["val total loss"])
    first = eval_loss()
    for _ in range(8):
        opt.zero_grad()
        total, _, _ = model.training_step(batch(), t_int=t_int, eps=eps)
        total.backward()
        opt.step()
    last = eval_loss()
    assert np.isfinite(last) and last < 0.9 * first, (first, last)
    # differentiable path (dropout off in eval -> use p = 0 by toggling) vs fused kernels on the updated weights
    for conv in model.dynamics.noise_predictor.conv_layers:
        conv.dropout.feat_dropout.p = 0.0
    total, _, _ = model.training_step(batch(), t_int=t_int, eps=eps)
    assert abs(float(total) - last) <= 2e-4 * max(1.0, abs(last)), (float(total), last)


def test_dataset_items_collate_into_a_training_batch(tmp_path, sd, dyn_cfg):
    """io.ProteinPharmacophoreDataset -> collate -> training_step (the train.py data path without DGL)."""
    from pharmacoforge_b200 import io as pio
    rng = np.random.default_rng(1)
    d = tmp_path / "split_0"
    d.mkdir()
    prot_n, ph_n, rp_n = np.array([60, 45, 80]), np.array([5, 4, 7]), np.array([3, 3, 3])
    mk = lambda cnt: np.stack([np.cumsum(cnt) - cnt, np.cumsum(cnt)], axis=1)
    np.savez(d / "prot_pharm_tensors.npz", prot_pos=rng.normal(size=(prot_n.sum(), 3)) * 6,
             prot_feat=rng.integers(0, 4, prot_n.sum()), pharm_pos=rng.normal(size=(ph_n.sum(), 3)) * 3,
             pharm_feat=rng.integers(0, 6, ph_n.sum()), prot_ph_pos=rng.normal(size=(rp_n.sum(), 3)),
             prot_ph_feat=rng.integers(0, 6, rp_n.sum()), prot_idx=mk(prot_n), pharm_idx=mk(ph_n), prot_ph_idx=mk(rp_n))
    ds = pio.ProteinPharmacophoreDataset([0], tmp_path, ['C', 'N', 'O', 'S', 'P', 'F', 'Cl', 'Br', 'I', 'B', 'D'])
    model = _model(sd, dyn_cfg, dropout=0.1).train()
    g = pio.ProteinPharmacophoreDataset.collate([ds[i] for i in range(3)], model, device="cuda:0")
    assert g.n_graphs == 3 and g.n_pharm == 16 and g.n_prot == 185
    total, losses, metrics = model.training_step(g)
    total.backward()
    assert np.isfinite(float(total)) and model.dynamics.pharm_encoder[0].weight.grad is not None


def test_fused_host_graph_matches_op_by_op_graph(sd, dyn_cfg):
    """train_fused.py (one autograd node per message chain / node update / noise-head stack) against the op-by-op graph
    of train_graph.py on the same batch, weights, (t, eps) and dropout masks (same seed -> same draws in the same order):
    same kernels in the same order, so losses are equal and gradients differ only by the float atomics of the
    LayerNorm / scatter backward kernels."""
    from pharmacoforge_b200 import train_graph
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    pockets = [Pocket.from_numpy(*make_pocket(120 + 30 * i, seed=70 + i)) for i in range(3)]
    sizes = [[5], [4], [7]]
    nf = sum(s[0] for s in sizes)
    gen = torch.Generator().manual_seed(11)
    h0 = torch.nn.functional.one_hot(torch.randint(0, 6, (nf,), generator=gen), 6).float()
    x0 = torch.cat([p.prot_x.mean(0, keepdim=True) + torch.randn(s[0], 3, generator=gen) for p, s in zip(pockets, sizes)])
    t_int = torch.tensor([15, 55, 90])
    eps = {"x": torch.randn(nf, 3, generator=gen), "h": torch.randn(nf, 6, generator=gen)}
    model = _model(sd, dyn_cfg, dropout=0.2).train()
    res = []
    was = train_graph.HOST_FUSED
    try:
        for fused in (True, False):
            train_graph.HOST_FUSED = fused
            model.zero_grad(set_to_none=True)
            torch.manual_seed(123)
            gb = GraphBatch.from_pockets(pockets, sizes, "cuda:0").set_pharmacophores(x0, h0)
            total, _, _ = model.training_step(gb, t_int=t_int, eps=eps)
            total.backward()
            res.append((float(total), {n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}))
    finally:
        train_graph.HOST_FUSED = was
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * max(1.0, abs(res[1][0])), (res[0][0], res[1][0])
    n_live = 0
    for n, ga in res[0][1].items():
        gb_ = res[1][1][n]
        assert (ga is None) == (gb_ is None), n
        if ga is not None and ga.numel():
            n_live += 1
            close(ga, gb_, rtol=1e-4, what=n)
    assert n_live >= 190


def test_training_dead_work_elimination_is_exact(sd, dyn_cfg):
    """dynamics.skip_dead_work in training mode: the last layer's protein side (pp / fp messages, protein update) is not
    run; losses and the gradient of every live parameter are the same numbers (same kernels on the same inputs for
    everything that is read), the dead parameters keep grad None either way."""
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    pockets = [Pocket.from_numpy(*make_pocket(140 + 25 * i, seed=80 + i)) for i in range(3)]
    sizes = [[6], [4], [8]]
    nf = sum(s[0] for s in sizes)
    gen = torch.Generator().manual_seed(21)
    h0 = torch.nn.functional.one_hot(torch.randint(0, 6, (nf,), generator=gen), 6).float()
    x0 = torch.cat([p.prot_x.mean(0, keepdim=True) + torch.randn(s[0], 3, generator=gen) for p, s in zip(pockets, sizes)])
    t_int = torch.tensor([5, 60, 95])
    eps = {"x": torch.randn(nf, 3, generator=gen), "h": torch.randn(nf, 6, generator=gen)}
    model = _model(sd, dyn_cfg, dropout=0.1).train()
    res = []
    try:
        for skip in (False, True):
            model.dynamics.skip_dead_work = skip
            model.zero_grad(set_to_none=True)
            torch.manual_seed(77)
            gb = GraphBatch.from_pockets(pockets, sizes, "cuda:0").set_pharmacophores(x0, h0)
            total, _, _ = model.training_step(gb, t_int=t_int, eps=eps)
            total.backward()
            res.append((float(total), {n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}))
    finally:
        model.dynamics.skip_dead_work = False
    # dropout masks: the skipped protein update does not draw its four masks, and it is the LAST draw of the forward pass, so
    # every mask that is used is the same in both runs
    assert abs(res[0][0] - res[1][0]) <= 1e-6 * max(1.0, abs(res[0][0])), (res[0][0], res[1][0])
    for n, ga in res[0][1].items():
        gb_ = res[1][1][n]
        assert (ga is None) == (gb_ is None), n
        if ga is not None and ga.numel():
            close(ga, gb_, rtol=1e-4, what=n)


def test_training_gradients_are_bit_reproducible(sd, dyn_cfg):
    """No floating-point atomics on the training path: split-K weight gradients, bias column sums and LayerNorm affine
    gradients are reduced by the last CTA in split order (the training workspace of the header), the gather's backward adds
    a row's edges in edge order.  Two backward passes from the same state give BIT-identical gradients on every parameter
    (64 pockets: the shapes of bench_train.py, where every one of those reductions is split over many CTAs), and the
    workspace's ticket counters are back at zero afterwards."""
    from pharmacoforge_b200 import train_ops
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    rng = np.random.default_rng(5)
    pockets = [Pocket.from_numpy(*make_pocket(int(rng.integers(250, 600)), seed=100 + i)) for i in range(24)]
    sizes = [[int(rng.integers(4, 9))] for _ in pockets]
    nf = sum(s[0] for s in sizes)
    gen = torch.Generator().manual_seed(9)
    x0 = torch.randn(nf, 3, generator=gen) * 4.0
    h0 = torch.nn.functional.one_hot(torch.randint(0, 6, (nf,), generator=gen), 6).float()
    t_int = torch.randint(1, 100, (len(pockets),), generator=gen)
    eps = {"x": torch.randn(nf, 3, generator=gen), "h": torch.randn(nf, 6, generator=gen)}
    model = _model(sd, dyn_cfg, dropout=0.0).train()
    grads = []
    for _ in range(2):
        gb = GraphBatch.from_pockets(pockets, sizes, "cuda:0")
        gb.set_pharmacophores(x0, h0)
        model.zero_grad(set_to_none=True)
        total, _, _ = model.training_step(gb, t_int=t_int, eps=eps)
        total.backward()
        grads.append({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    assert len(grads[0]) >= 190 and grads[0].keys() == grads[1].keys()
    for n in grads[0]:
        assert torch.isfinite(grads[0][n]).all(), n
        assert torch.equal(grads[0][n], grads[1][n]), n
    ws = train_ops._workspace(torch.device("cuda:0"))
    tail = _lib_tail_words()
    assert int(ws.view(torch.int32)[-tail:].abs().sum()) == 0


def _lib_tail_words():
    import re
    import os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pharmacoforge_b200.h")).read()
    return int(re.search(r"#define PF_TRAIN_WS_TAIL (\d+)", hdr).group(1)) // 4
