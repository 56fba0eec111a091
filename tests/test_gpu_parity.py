"""GPU parity tests: every CUDA entry point (through the torch.library ops over the C ABI) against the CPU
oracle on the same seeded inputs, and against the golden fixtures produced by the reference's own code.

Tolerances: edge lists / indices / tile plans bit-exact; the posterior + COM step bit-exact given identical eps;
floating-point features within rtol 1e-4 (+ atol 1e-5 for entries near zero), the fp32 bar BASELINE.json states.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-5


def t(a):
    return torch.from_numpy(np.asarray(a))


def canon(src, dst):
    s = src.detach().cpu().numpy().astype(np.int64)
    d = dst.detach().cpu().numpy().astype(np.int64)
    o = np.lexsort((s, d))
    return s[o], d[o]


def close(a, b, rtol=RTOL, atol=ATOL, what=""):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    bound = atol + rtol * np.abs(b)
    assert (err <= bound).all(), f"{what}: max abs err {err.max():.3e}, worst ratio {(err / bound).max():.2f}"


# Tolerance of the single-pass fp16 mode (PF_FLAG_FP16_SINGLE_PASS, BASELINE.json configs[3] "bf16 edge-MLP path"):
# max |got - want| <= FP16_TOL * max |want| per tensor.  11-bit operands and a 2^-11 tanh: ~1e-3 per contraction.
FP16_TOL = 2e-2


def within(a, b, frac, what=""):
    a = a.detach().cpu().double() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.detach().cpu().double() if torch.is_tensor(b) else torch.as_tensor(b).double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ok = torch.isfinite(a).all()
    err, scale = (a - b).abs().max().item(), b.abs().max().item()
    assert ok and err <= frac * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (allowed {frac:g} x scale)"
    return err / max(scale, 1e-30)


def to_cm(v):   # reference vector layout [N,16,3] -> kernel layout [N,3,16] flattened to [N,48]
    return v.permute(0, 2, 1).reshape(v.shape[0], 48).contiguous()


def from_cm(v):
    return v.reshape(v.shape[0], 3, 16).permute(0, 2, 1).contiguous()


@pytest.fixture(scope="module")
def env(sd, dyn_cfg):
    import pf_oracle as O
    from pharmacoforge_b200 import ops
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    from pharmacoforge_b200.synthetic import make_pocket
    cfg = dict(dyn_cfg)
    gcut = cfg.pop("graph_cutoffs")
    model = PharmacophoreDiff(6, 11, ["Aromatic", "HydrogenDonor", "HydrogenAcceptor", "PositiveIon", "NegativeIon",
                                      "Hydrophobic"], n_timesteps=100, graph_config={"graph_cutoffs": gcut},
                              dynamics_config=cfg, precision=1e-5)
    model.load_state_dict(sd, strict=True)
    model.eval()

    class Env:
        pass
    e = Env()
    e.O, e.ops, e.model, e.sd, e.cfg = O, ops, model, sd, dyn_cfg
    e.GraphBatch, e.Pocket, e.make_pocket = GraphBatch, Pocket, make_pocket
    e.dev = torch.device("cuda:0")
    e.W = model.dynamics.packed_weights(e.dev)

    def build(pocket_specs, sizes):
        """pocket_specs: [(n_atoms, seed)]; returns (GraphBatch, oracle FlatBatch)"""
        pk = [make_pocket(n, seed=s) for n, s in pocket_specs]
        g = GraphBatch.from_pockets([Pocket.from_numpy(p, h) for p, h in pk], sizes, e.dev)
        b = O.build_batch([(t(p), t(h)) for p, h in pk], sizes)
        return g, b
    e.build = build

    def set_state(g, b, pharm_x, pharm_h, prot_x=None):
        st = model.dynamics.bind(g)
        g.pharm_x.copy_(pharm_x)
        g.pharm_h.copy_(pharm_h)
        b.pharm_x, b.pharm_h = pharm_x.clone(), pharm_h.clone()
        if prot_x is not None:
            g.prot_x.copy_(prot_x)
            b.prot_x = prot_x.clone()
        return st
    e.set_state = set_state
    return e


def random_state(b, seed, spread=3.0):
    gen = torch.Generator().manual_seed(seed)
    nf = int(b.pharm_ptr[-1])
    x = torch.randn(nf, 3, generator=gen) * spread
    h = torch.randn(nf, 6, generator=gen)
    # centre every graph's protein near its pharmacophore, as the sampler's frame does
    com = torch.stack([b.prot_x[int(b.prot_ptr[i]):int(b.prot_ptr[i + 1])].mean(0) for i in range(b.n_graphs)])
    prot = b.prot_x - com[b.prot_b] + torch.randn(b.n_graphs, 3, generator=gen)[b.prot_b]
    return x, h, prot


# ------------------------------------------------------------------------------------------------ K1
def test_pp_radius_csr_matches_golden_and_oracle(env, golden):
    g, b = env.build([(400, 0)], [[3]])
    src, dst = canon(*g.pp_edges())
    gold = golden("pp_graph_n400_seed0.npz")
    assert np.array_equal(src, gold["src"]) and np.array_equal(dst, gold["dst"])
    assert np.array_equal(g.pp_rowptr.cpu().numpy()[1:] - g.pp_rowptr.cpu().numpy()[:-1], g.pp_cnt.cpu().numpy())


def test_pp_radius_csr_ragged_batch(env):
    g, b = env.build([(400, 1), (37, 2), (250, 3), (1, 4), (1500, 5)], [[3, 8], [4], [5, 6, 7], [3], [16]])
    s, d = canon(*g.pp_edges())
    so, do = canon(*b.pp)
    assert np.array_equal(s, so) and np.array_equal(d, do)
    g.check_status()


def test_radius_max_neighbors_truncation(env):
    # dense cluster: more neighbours than max_num_neighbors -> first `max` in ascending index
    O, ops = env.O, env.ops
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(300, 3, generator=gen) * 4.0
    ptr = torch.tensor([0, 120, 300], dtype=torch.int32)
    rowptr, deg, col = ops.radius_csr(x.cuda(), ptr.cuda(), 3.0, 20)
    so, do = O.radius_edges(x, ptr.long(), 3.0, 20)
    dst = torch.repeat_interleave(torch.arange(300), deg.cpu().long())
    s, d = canon(col.cpu(), dst)
    so, do = canon(so, do)
    assert deg.max().item() <= 21 and np.array_equal(s, so) and np.array_equal(d, do)


@pytest.mark.parametrize("case", ["pockets", "dense_truncated", "very_dense_fallback", "sparse_far_apart", "degenerate"])
def test_cell_list_radius_csr_is_bit_identical_to_all_pairs(env, case):
    """K1 as a cell list (pf_cell_radius_count / fill) against the all-pairs kernel and the oracle: same CSR bit for bit,
    including torch_cluster's truncation rule, rows denser than the in-kernel sort buffer (ordered-scan fallback),
    point clouds so sparse that the cell edge has to grow, and empty / single-atom / coincident-point segments."""
    O, ops = env.O, env.ops
    gen = torch.Generator().manual_seed(3)
    r, mx = 3.5, 100
    if case == "pockets":
        xs = [t(env.make_pocket(n, seed=s)[0]) for n, s in ((400, 0), (37, 2), (1, 4), (1500, 5), (250, 3))]
    elif case == "dense_truncated":
        xs, r, mx = [torch.rand(120, 3, generator=gen) * 4.0, torch.rand(180, 3, generator=gen) * 4.0 + 30.0], 3.0, 20
    elif case == "very_dense_fallback":
        xs, mx = [torch.rand(300, 3, generator=gen) * 1.5 - 40.0, torch.rand(50, 3, generator=gen) * 2.0], 100
    elif case == "sparse_far_apart":
        xs = [torch.rand(40, 3, generator=gen) * 5000.0, torch.cat([torch.rand(20, 3, generator=gen) * 3.0,
                                                                   torch.rand(20, 3, generator=gen) * 3.0 + 1.0e4])]
    else:
        xs = [torch.zeros(0, 3), torch.ones(1, 3) * 7.0, torch.ones(6, 3) * -2.5, torch.rand(9, 3, generator=gen)]
    x = torch.cat(xs).float().contiguous()
    ptr = torch.tensor(np.concatenate([[0], np.cumsum([v.shape[0] for v in xs])]), dtype=torch.int32)
    r0, d0, c0 = ops.radius_csr(x.cuda(), ptr.cuda(), r, mx)
    r1, d1, c1 = ops.cell_radius_csr(x.cuda(), ptr.cuda(), r, mx)
    assert torch.equal(r0, r1) and torch.equal(d0, d1) and torch.equal(c0, c1), case
    so, do = O.radius_edges(x, ptr.long(), r, mx)
    dst = torch.repeat_interleave(torch.arange(x.shape[0]), d1.cpu().long())
    s, d = canon(c1.cpu(), dst)
    so, do = canon(so, do)
    assert np.array_equal(s, so) and np.array_equal(d, do), case
    # rows come out sorted by source index (the reference order of torch_cluster.radius_graph within a destination)
    c = c1.cpu().numpy()
    rp = r1.cpu().numpy()
    assert all(np.all(np.diff(c[rp[i]:rp[i + 1]]) > 0) for i in range(x.shape[0]))


def test_pp_graph_built_per_pocket_then_replicated(env, monkeypatch):
    """GraphBatch builds the pp CSR once per DISTINCT pocket and replicates it per graph on the device
    (protein_pharm_dataset.py:234-236 + copy_graph): identical, bit for bit, to the all-pairs build over the replicated
    batch (PF_K1=brute), also when a graph_range cuts through a pocket's samples."""
    pk = [env.make_pocket(n, seed=s) for n, s in ((400, 1), (37, 2), (1, 4), (250, 3))]
    pockets = [env.Pocket.from_numpy(p, h) for p, h in pk]
    sizes = [[3, 8, 5], [4], [3, 3], [5, 6, 7]]
    for rng in (None, range(2, 8)):
        g1 = env.GraphBatch.from_pockets(pockets, sizes, env.dev, graph_range=rng)
        monkeypatch.setenv("PF_K1", "brute")
        g0 = env.GraphBatch.from_pockets(pockets, sizes, env.dev, graph_range=rng)
        monkeypatch.delenv("PF_K1")
        for name in ("pp_rowptr", "pp_cnt", "pp_col", "pp_tiles", "pp_n_tiles"):
            assert torch.equal(getattr(g0, name), getattr(g1, name)), (name, rng)
        assert g0.n_pp_edges == g1.n_pp_edges and g0.pp_num_tiles == g1.pp_num_tiles


def test_ordered_tile_plan_matches_atomic_plan(env):
    """The static pp plan lists its tiles in graph order (L2 reuse of source rows across the tiles of a graph): same tiles
    as the atomic planner, sorted; every tile inside one graph, <= 128 rows, whole destinations only."""
    g, b = env.build([(400, 1), (37, 2), (1, 4), (250, 3), (1500, 5)], [[3, 8, 5], [4], [3], [5, 6, 7], [16]])
    n = g.pp_num_tiles
    tiles = g.pp_tiles.cpu().numpy().reshape(-1, 2)[:n]
    assert np.all(tiles[1:, 0] >= tiles[:-1, 1]) and tiles[0, 0] == 0 and tiles[-1, 1] == g.n_prot   # ordered, covering
    t2 = torch.zeros_like(g.pp_tiles)
    n2 = torch.zeros(1, dtype=torch.int32, device=env.dev)
    env.ops.plan_tiles(g.pp_cnt, g.prot_ptr, False, 128, t2, n2, g.status)
    other = t2.cpu().numpy().reshape(-1, 2)[:int(n2)]
    assert int(n2) == n and np.array_equal(other[np.argsort(other[:, 0])], tiles)
    cnt = g.pp_cnt.cpu().numpy()
    ptr = g.prot_ptr_host
    for s0, s1 in tiles:
        assert cnt[s0:s1].sum() <= 128 and s1 - s0 <= 128
        assert np.searchsorted(ptr, s0, side="right") == np.searchsorted(ptr, s1 - 1, side="right")   # one graph


def test_exclusive_scan(env):
    for n in (0, 1, 5, 2048, 2049, 100_003, 3_000_017):
        x = torch.randint(0, 20, (n,), dtype=torch.int32)
        out = env.ops.exclusive_scan(x.cuda()).cpu()
        ref = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(x.long(), 0)])
        assert torch.equal(out.long(), ref), n


# ------------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("spread", [1.0, 4.0, 12.0])
def test_dynamic_graph_bit_exact(env, spread):
    g, b = env.build([(400, 0), (60, 7), (3, 9)], [[3, 4, 5, 6, 7, 8], [16, 1], [4]])
    x, h, prot = random_state(b, 11, spread)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    ops = env.ops
    ops.dyn_graph(g.prot_x, g.prot_ptr, g.pharm_x, g.pharm_ptr, 9.0, 200, 5, g.ff_start, g.ff_cnt, g.ff_col, g.pf_cnt,
                  g.pf_col, g.fp_seg_dst, g.fp_seg_start, g.fp_seg_cnt, g.fp_col, g.status)
    g.check_status()
    got = g.dynamic_edges()
    want = env.O.dynamic_edges(b, 9.0, 5)
    for et in ("ff", "pf", "fp"):
        s, d = canon(*got[et])
        so, do = canon(*want[et])
        assert np.array_equal(s, so) and np.array_equal(d, do), et
    # pf rows are in ascending-distance order with ties to the lower index, exactly like knn()
    q, c = env.O.knn_edges(b.prot_x, b.prot_ptr, b.pharm_x, b.pharm_ptr, 5)
    assert np.array_equal(got["pf"][0].cpu().numpy(), c.numpy()) and np.array_equal(got["pf"][1].cpu().numpy(), q.numpy())


@pytest.mark.parametrize("ff_k", [1, 3, 20])
def test_ff_knn_graph_variant_bit_exact(env, ff_k):
    """dynamics.ff_k > 0 (dynamics_gvp.py:193-194): ff edges = knn_graph(pharm x_t, k = ff_k).  Edge list bit-exact against the
    oracle (graphs smaller than k + 1 keep all their other nodes; single-node graphs have no edge), pf / fp unchanged."""
    g, b = env.build([(400, 0), (60, 7), (3, 9), (30, 2)], [[3, 4, 5, 6, 7, 8], [16, 1], [4], [2, 12]])
    x, h, prot = random_state(b, 11, 2.0)
    x[5] = x[4]                                   # coincident centres: the (distance, index) tie rule
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    env.ops.dyn_graph(g.prot_x, g.prot_ptr, g.pharm_x, g.pharm_ptr, 9.0, 200, 5, g.ff_start, g.ff_cnt, g.ff_col, g.pf_cnt,
                      g.pf_col, g.fp_seg_dst, g.fp_seg_start, g.fp_seg_cnt, g.fp_col, g.status, ff_k)
    g.check_status()
    want = env.O.dynamic_edges(b, 9.0, 5, ff_k)
    got = g.dynamic_edges()
    for et in ("ff", "pf", "fp"):
        s, d = canon(*got[et])
        so, do = canon(*want[et])
        assert np.array_equal(s, so) and np.array_equal(d, do), et
    nf = np.diff(g.pharm_ptr_host)
    assert got["ff"][0].numel() == int(sum(n * min(ff_k, n - 1) for n in nf))


def test_denoiser_with_ff_knn_graph(env):
    """The whole denoiser under dynamics_config ff_k = 4 against the oracle with the same switch."""
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    cfg = dict(env.cfg, ff_k=4)
    gcut = cfg.pop("graph_cutoffs")
    model = PharmacophoreDiff(6, 11, list("abcdef"), n_timesteps=100, graph_config={"graph_cutoffs": gcut},
                              dynamics_config=cfg, precision=1e-5)
    model.load_state_dict(env.sd, strict=True)
    model.eval()
    g, b = env.build([(300, 5), (120, 6)], [[8, 3, 6], [5, 7]])
    x, h, prot = random_state(b, 23, 3.0)
    g.pharm_h = h.cuda().clone()
    g.pharm_x.copy_(x.cuda())
    g.prot_x.copy_(prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    tt = torch.tensor([0.2, 0.2, 0.9, 0.5, 0.61])
    wh, wx = env.O.denoiser(env.sd, b, tt, dict(cfg, graph_cutoffs=gcut))
    gh, gx = model.dynamics(g, tt, None)
    close(gh, wh, what="eps_h (ff kNN)")
    close(gx, wx, what="eps_x (ff kNN)")
    n_knn = int(g.ff_cnt.sum().item())
    assert n_knn == env.O.dynamic_edges(b, 9.0, 5, 4)["ff"][0].numel() < env.O.dynamic_edges(b, 9.0, 5, 0)["ff"][0].numel()
    # the differentiable training graph builds the same kNN ff edges (it used to ignore ff_k) and gives the same eps
    from pharmacoforge_b200 import train_graph
    g.pharm_h.copy_(h.cuda())
    g.pharm_x.copy_(x.cuda())
    g.prot_x.copy_(prot.cuda())
    model.cuda()          # the training graph reads the nn.Parameters themselves (the fused path packs its own device images)
    th, tx = train_graph.dynamics_forward(model.dynamics, g, tt, training=False)
    assert int(g.ff_cnt.sum().item()) == n_knn
    close(th.detach(), wh, rtol=2e-4, what="train eps_h (ff kNN)")
    close(tx.detach(), wx, rtol=2e-4, what="train eps_x (ff kNN)")


def test_knn_ties_prefer_lower_index(env):
    # duplicate protein atoms -> exact distance ties
    O = env.O
    pos = np.zeros((8, 3), dtype=np.float32)
    pos[:, 0] = [1, 1, 2, 2, 3, 3, 4, 4]
    onehot = np.zeros((8, 11), dtype=np.float32)
    onehot[:, 0] = 1
    g = env.GraphBatch.from_pockets([env.Pocket.from_numpy(pos, onehot)], [[2]], env.dev)
    b = O.build_batch([(t(pos), t(onehot))], [[2]])
    x = torch.tensor([[0.0, 0, 0], [5.0, 0, 0]])
    env.set_state(g, b, x.cuda(), torch.zeros(2, 6).cuda())
    env.ops.dyn_graph(g.prot_x, g.prot_ptr, g.pharm_x, g.pharm_ptr, 9.0, 200, 5, g.ff_start, g.ff_cnt, g.ff_col,
                      g.pf_cnt, g.pf_col, g.fp_seg_dst, g.fp_seg_start, g.fp_seg_cnt, g.fp_col, g.status)
    q, c = O.knn_edges(b.prot_x, b.prot_ptr, x, b.pharm_ptr, 5)
    assert np.array_equal(g.pf_col.cpu().numpy()[:10], c.numpy())


# ------------------------------------------------------------------------------------------------ K0 / K3 / K4 / K5a
def test_encoders(env):
    g, b = env.build([(120, 3)], [[3, 5, 8, 6]])
    x, h, prot = random_state(b, 5)
    tt = torch.tensor([0.37, 0.99, 0.01, 0.5])
    got_f = env.ops.encode(h.cuda(), g.pharm_ptr, tt.cuda(), env.W.view("pharm_enc"))
    got_p = env.ops.encode(g.prot_feats, g.prot_ptr, tt.cuda(), env.W.view("prot_enc"))
    close(got_f, env.O.encoder(env.sd, "dynamics.pharm_encoder", h, tt[b.pharm_b]), what="pharm enc")
    close(got_p, env.O.encoder(env.sd, "dynamics.prot_encoder", b.prot_h, tt[b.prot_b]), what="prot enc")


def test_encoder_one_hot_table_is_bit_identical(env):
    """pf_encode computes the nf distinct rows of a one-hot graph once per graph: same bits as the per-node path (small
    graphs take it), and rows that are not one-hot fall back to it inside a large graph."""
    gen = torch.Generator().manual_seed(12)
    n, nf = 240, 11
    feats = torch.nn.functional.one_hot(torch.randint(0, nf, (n,), generator=gen), nf).float()
    feats[7] = torch.randn(nf, generator=gen)          # not one-hot
    feats[8] = 0.0                                      # all zero
    feats[9, :] = 0.0
    feats[9, 2] = 1.0
    feats[9, 5] = 1.0                                   # two ones
    feats[10] = feats[10] * 2.0                         # a single 2.0
    w = env.W.view("prot_enc")
    t_big = torch.tensor([0.63])
    big = env.ops.encode(feats.cuda(), torch.tensor([0, n], dtype=torch.int32).cuda(), t_big.cuda(), w)
    ptr_small = torch.arange(0, n + 1, 10, dtype=torch.int32)     # 24 graphs of 10 nodes: below the table threshold
    small = env.ops.encode(feats.cuda(), ptr_small.cuda(), t_big.repeat(24).cuda(), w)
    assert torch.equal(big, small)
    close(big, env.O.encoder(env.sd, "dynamics.prot_encoder", feats, t_big.repeat(n)), what="prot enc (table path)")


def _edge_conv_case(env, etype_idx, layer, with_vectors, impl="tc", fp16=False):
    O, ops = env.O, env.ops
    g, b = env.build([(150, 3), (90, 4)], [[3, 5, 8], [6, 4]])
    x, h, prot = random_state(b, 21)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    ops.dyn_graph(g.prot_x, g.prot_ptr, g.pharm_x, g.pharm_ptr, 9.0, 200, 5, g.ff_start, g.ff_cnt, g.ff_col, g.pf_cnt,
                  g.pf_col, g.fp_seg_dst, g.fp_seg_start, g.fp_seg_cnt, g.fp_col, g.status)
    edges = O.dynamic_edges(b, 9.0, 5)
    gen = torch.Generator().manual_seed(3)
    feats = {"pharm": (torch.randn(g.n_pharm, 128, generator=gen), x,
                       torch.randn(g.n_pharm, 16, 3, generator=gen) * (1.0 if with_vectors else 0.0)),
             "prot": (torch.randn(g.n_prot, 128, generator=gen), prot,
                      torch.randn(g.n_prot, 16, 3, generator=gen) * (1.0 if with_vectors else 0.0))}
    snt, et, dnt = O.ETYPES[etype_idx]
    src, dst = edges[et]
    key = f"dynamics.noise_predictor.conv_layers.{layer}.edge_message_fns.{snt}_{et}_{dnt}"
    ms, mv = O.edge_messages(env.sd, key, feats[snt][0][src], feats[snt][2][src], feats[snt][1][src],
                             feats[dnt][1][dst])
    n_dst = feats[dnt][0].shape[0]
    want_h, want_v = O.mean_aggregate(ms, dst, n_dst), O.mean_aggregate(mv, dst, n_dst)
    seg = {"ff": (g.ff_start, g.ff_cnt, None, g.ff_col, g.pharm_chunk_ptr),
           "pf": (g.pf_start, g.pf_cnt, None, g.pf_col, g.pharm_chunk_ptr),
           "fp": (g.fp_seg_start, g.fp_seg_cnt, g.fp_seg_dst, g.fp_col, g.fp_chunk_ptr),
           "pp": (g.pp_start, g.pp_cnt, None, g.pp_col, g.prot_ptr)}[et]
    tiles = torch.zeros(2 * (g.n_prot + g.dyn_max_tiles), dtype=torch.int32, device=env.dev)
    n_tiles = torch.zeros(1, dtype=torch.int32, device=env.dev)
    accumulate = et in ("pf", "fp")
    ops.plan_tiles(seg[1], seg[4], accumulate, 128 if impl == "tc" else 64, tiles, n_tiles, g.status)
    base_h = torch.randn(n_dst, 128, generator=gen) if accumulate else torch.full((n_dst, 128), float("nan"))
    base_v = torch.randn(n_dst, 48, generator=gen) if accumulate else torch.full((n_dst, 48), float("nan"))
    agg_h, agg_v = base_h.cuda(), base_v.cuda()
    src_v = to_cm(feats[snt][2]).cuda() if with_vectors else None
    if impl == "tc":
        blob = env.W.tc[(layer * 4 + etype_idx) * env.W.tc_stride:(layer * 4 + etype_idx + 1) * env.W.tc_stride]
        ops.edge_conv_tc(feats[snt][0].cuda(), src_v, feats[snt][1].cuda().contiguous(),
                         feats[dnt][1].cuda().contiguous(), seg[0], seg[1], seg[2], seg[3], tiles, n_tiles, blob, agg_h,
                         agg_v, accumulate, fp16)
    else:
        ops.edge_conv(feats[snt][0].cuda(), src_v, feats[snt][1].cuda().contiguous(),
                      feats[dnt][1].cuda().contiguous(), seg[0], seg[1], seg[2], seg[3], tiles, n_tiles,
                      env.W.view(f"msg{layer}_{etype_idx}"), 3, agg_h, agg_v, accumulate)
    torch.cuda.synchronize()
    g.check_status()
    if accumulate:
        want_h, want_v = base_h + want_h, base_v + to_cm(want_v)
    else:
        want_v = to_cm(want_v)
    if fp16:
        within(agg_h, want_h, FP16_TOL, what=f"{et} scalars (fp16 single pass)")
        within(agg_v, want_v, FP16_TOL, what=f"{et} vectors (fp16 single pass)")
        return
    close(agg_h, want_h, what=f"{et} scalars")
    close(agg_v, want_v, what=f"{et} vectors")


@pytest.mark.parametrize("impl", ["tc", "ffma"])
@pytest.mark.parametrize("etype_idx", [0, 1, 2, 3])
@pytest.mark.parametrize("with_vectors", [False, True])
def test_edge_conv_each_etype(env, etype_idx, with_vectors, impl):
    _edge_conv_case(env, etype_idx, 1 if with_vectors else 0, with_vectors, impl)


@pytest.mark.parametrize("fp16", [False, True])
def test_edge_conv_seeded_first_layer(env, fp16):
    """pf_seed_table + pf_edge_conv_tc_seeded (first conv layer, one-hot protein features): the per-node part of GVP 0 comes
    from the (graph, atom type) table.  Checked against the oracle's edge messages computed from the full per-edge input
    [h_src; rbf; sh] (gvp.py:540-551), and against the general kernel on the same inputs."""
    O, ops = env.O, env.ops
    g, b = env.build([(150, 3), (90, 4), (1, 9)], [[3, 5, 8], [6, 4], [2]])
    x, h, prot = random_state(b, 5)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    seed_row, seed_rep = g.seed_arrays()
    # table rows are (graph, type); a graph's atoms of one type share their encoder row: build h that way
    tt = torch.tensor([0.1, 0.55, 0.9])
    enc = ops.encode(g.prot_feats, g.prot_ptr, tt.cuda(), env.W.view("prot_enc"))
    rows = seed_row.long().cpu()
    assert int(rows.max()) < seed_rep.numel() and torch.equal(enc.cpu()[seed_rep.long().cpu()[rows]], enc.cpu())
    table = torch.full((seed_rep.numel(), 128), float("nan"), device=env.dev)
    ops.seed_table(enc, seed_rep, env.W.view("msg0_3"), table)
    Wf = env.sd["dynamics.noise_predictor.conv_layers.0.edge_message_fns.prot_pp_prot.0.to_feats_out.0.weight"]
    used = seed_rep.cpu() >= 0
    want_tab = -1.4426950408889634 * enc.cpu().double()[seed_rep.long().cpu()[used]] @ Wf[:, :128].double().t()
    close(table.cpu()[used], want_tab, rtol=1e-5, atol=1e-6, what="seed table")
    assert torch.isnan(table.cpu()[~used]).all()          # rows without a representative are never written
    src, dst = (v.cpu() for v in g.pp_edges())
    key = "dynamics.noise_predictor.conv_layers.0.edge_message_fns.prot_pp_prot"
    zeros = torch.zeros(g.n_prot, 16, 3)
    ms, mv = O.edge_messages(env.sd, key, enc.cpu()[src], zeros[src], prot[src], prot[dst])
    want_h, want_v = O.mean_aggregate(ms, dst, g.n_prot), to_cm(O.mean_aggregate(mv, dst, g.n_prot))
    blob = env.W.tc[3 * env.W.tc_stride:4 * env.W.tc_stride]
    out = {}
    for seeded in (True, False):
        agg_h = torch.full((g.n_prot, 128), float("nan"), device=env.dev)
        agg_v = torch.full((g.n_prot, 48), float("nan"), device=env.dev)
        if seeded:
            ops.edge_conv_tc_seeded(seed_row, table, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles,
                                    g.pp_n_tiles, blob, agg_h, agg_v, False, fp16)
        else:
            ops.edge_conv_tc(enc, None, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles,
                             g.pp_n_tiles, blob, agg_h, agg_v, False, fp16)
        torch.cuda.synchronize()
        g.check_status()
        out[seeded] = (agg_h.cpu(), agg_v.cpu())
        if fp16:
            within(agg_h, want_h, FP16_TOL, what="seeded pp scalars (fp16)")
            within(agg_v, want_v, FP16_TOL, what="seeded pp vectors (fp16)")
        else:
            close(agg_h, want_h, what=f"pp scalars (seeded={seeded})")
            close(agg_v, want_v, what=f"pp vectors (seeded={seeded})")
    if not fp16:   # the two kernels agree far inside the parity bar (they differ only in summation order of GVP 0)
        close(out[True][0], out[False][0], rtol=2e-5, atol=2e-6, what="seeded vs general scalars")
        close(out[True][1], out[False][1], rtol=2e-5, atol=2e-6, what="seeded vs general vectors")


def test_denoiser_layer0_encoder_table_is_bit_identical(env):
    """PF_FLAG_NO_LAYER0_TABLE: reading the first layer's protein scalars from the (graph, atom type) encoder table through
    the row map (pf_edge_conv_tc_mapped / pf_node_update_tc_mapped) is the same arithmetic on the same values as the
    per-node encoder pass: eps and the final protein features agree bit for bit, ragged batch, per-graph timesteps."""
    g, b = env.build([(400, 0), (250, 1), (1, 4), (77, 9)], [[3, 8, 5], [4, 6], [3], [16, 3]])
    x, h, prot = random_state(b, 23, 4.0)
    st = env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    tt = torch.rand(g.n_graphs, generator=torch.Generator().manual_seed(3))
    dyn = env.model.dynamics
    res = {}
    try:
        for on in (True, False):
            dyn.layer0_table = on
            st.prot_h.fill_(float("nan"))
            gh, gx = dyn(g, tt, None)
            res[on] = (gh.clone(), gx.clone(), st.prot_h.clone(), st.prot_v.clone())
    finally:
        dyn.layer0_table = True
    for a_, b_ in zip(res[True], res[False]):
        assert torch.equal(a_, b_)
    wh, wx = env.O.denoiser(env.sd, b, tt, env.cfg)
    close(res[True][0], wh, what="eps_h (encoder table)")
    close(res[True][1], wx, what="eps_x (encoder table)")


def test_denoiser_layer0_seed_switch(env):
    """PF_FLAG_NO_LAYER0_SEED: the general first-layer kernel and the seeded one give the same eps (per-graph timesteps);
    a batch whose protein features are not one-hot falls back to the general kernel by itself."""
    g, b = env.build([(400, 0), (250, 1)], [[3, 8, 5], [4, 6]])
    x, h, prot = random_state(b, 17, 4.0)
    st = env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    assert st.seed is not None and st.args.n_seed_rows == 5 * 11
    tt = torch.tensor([0.03, 0.5, 0.5, 0.77, 1.0])
    wh, wx = env.O.denoiser(env.sd, b, tt, env.cfg)
    dyn = env.model.dynamics
    res = {}
    try:
        for on in (True, False):
            dyn.layer0_seed = on
            gh, gx = dyn(g, tt, None)
            res[on] = (gh.clone(), gx.clone())
            close(gh, wh, what=f"eps_h (layer0_seed={on})")
            close(gx, wx, what=f"eps_x (layer0_seed={on})")
    finally:
        dyn.layer0_seed = True
    close(res[True][0], res[False][0], rtol=2e-5, atol=2e-6, what="eps_h seeded vs general")
    close(res[True][1], res[False][1], rtol=2e-5, atol=2e-6, what="eps_x seeded vs general")
    # soft (not one-hot) protein features: no seed arrays, general kernel, still correct
    g2, b2 = env.build([(120, 2)], [[4, 7]])
    soft = torch.softmax(torch.randn(g2.n_prot, 11, generator=torch.Generator().manual_seed(1)), dim=1)
    g2.prot_feats.copy_(soft.cuda())
    b2.prot_h = soft.clone()
    x, h, prot = random_state(b2, 3)
    st2 = env.set_state(g2, b2, x.cuda(), h.cuda(), prot.cuda())
    b2.pharm_x, b2.pharm_h, b2.prot_x = x, h, prot
    assert st2.seed is None
    tt2 = torch.tensor([0.3, 0.6])
    wh, wx = env.O.denoiser(env.sd, b2, tt2, env.cfg)
    gh, gx = dyn(g2, tt2, None)
    close(gh, wh, what="eps_h (soft features)")
    close(gx, wx, what="eps_x (soft features)")


@pytest.mark.parametrize("impl", ["tc", "ffma"])
def test_node_update(env, impl):
    O, ops = env.O, env.ops
    gen = torch.Generator().manual_seed(8)

    def run(h_in, v_in, agg_h, agg_v, layer, nt, h_out, v_out):
        if impl == "tc":
            ops.node_update_tc(h_in, v_in, agg_h, agg_v, env.W.tcu_view(layer, nt), h_out, v_out)
        else:
            ops.node_update(h_in, v_in, agg_h, agg_v, env.W.view(f"upd{layer}_{nt}"), 2, h_out, v_out)
        torch.cuda.synchronize()

    for n in (1, 63, 64, 65, 127, 128, 129, 1000, 40000):
        h, v = torch.randn(n, 128, generator=gen), torch.randn(n, 16, 3, generator=gen)
        ah, av = torch.randn(n, 128, generator=gen), torch.randn(n, 16, 3, generator=gen)
        p = "dynamics.noise_predictor.conv_layers.1"
        s, vv = O.gvp_layernorm(env.sd, f"{p}.message_layer_norms.prot", h + ah, v + av)
        rs, rv = s, vv
        for i in range(2):
            rs, rv = O.gvp(env.sd, f"{p}.node_update_fns.prot.{i}", rs, rv)
        want_h, want_v = O.gvp_layernorm(env.sd, f"{p}.update_layer_norms.prot", s + rs, vv + rv)
        hd, vd = h.cuda(), to_cm(v).cuda()
        run(hd, vd, ah.cuda(), to_cm(av).cuda(), 1, 1, hd, vd)   # in place
        close(hd, want_h, what=f"node_update h n={n}")
        close(from_cm(vd), want_v, what=f"node_update v n={n}")
    # zero input vectors (first layer): v_in = None
    h, ah, av = torch.randn(70, 128, generator=gen), torch.randn(70, 128, generator=gen), torch.randn(70, 16, 3, generator=gen)
    p = "dynamics.noise_predictor.conv_layers.0"
    s, vv = O.gvp_layernorm(env.sd, f"{p}.message_layer_norms.pharm", h + ah, av)
    rs, rv = s, vv
    for i in range(2):
        rs, rv = O.gvp(env.sd, f"{p}.node_update_fns.pharm.{i}", rs, rv)
    want_h, want_v = O.gvp_layernorm(env.sd, f"{p}.update_layer_norms.pharm", s + rs, vv + rv)
    ho, vo = torch.empty(70, 128, device=env.dev), torch.empty(70, 48, device=env.dev)
    run(h.cuda(), None, ah.cuda(), to_cm(av).cuda(), 0, 0, ho, vo)
    close(ho, want_h, what="node_update h (v=0)")
    close(from_cm(vo), want_v, what="node_update v (v=0)")


def test_noise_head(env):
    gen = torch.Generator().manual_seed(9)
    for n in (5, 64, 165):
        h, v = torch.randn(n, 128, generator=gen), torch.randn(n, 16, 3, generator=gen)
        wh, wx = env.O.noise_head(env.sd, "dynamics.noise_predictor.noise_predictor", h, v)
        gh, gx = env.ops.noise_head(h.cuda(), to_cm(v).cuda(), env.W.view("noise"), 4, 6)
        close(gh, wh, what="eps_h")
        close(gx, wx, what="eps_x")


# ------------------------------------------------------------------------------------------------ denoiser
def test_denoiser_call_against_reference_golden(env, golden):
    """The fixture holds inputs, edge lists, per-layer features and outputs of the REFERENCE's own code."""
    d = golden("denoiser_call.npz")
    sizes = [int(v) for v in d["sizes"]]
    g, b = env.build([(int(d["n_atoms"]), int(d["pocket_seed"]))], [sizes])
    st = env.set_state(g, b, t(d["x_t"]).cuda(), t(d["h_t"]).cuda(), t(d["prot_x"]).cuda())
    eps_h, eps_x = env.model.dynamics(g, t(d["t"]), None)
    g.check_status()
    got = g.dynamic_edges()
    got["pp"] = g.pp_edges()
    for et in ("ff", "pf", "fp", "pp"):
        s, dd = canon(*got[et])
        assert np.array_equal(s, d[f"e_{et}_src"]) and np.array_equal(dd, d[f"e_{et}_dst"]), et
    close(st.pharm_hh, d["conv1_pharm_h"], what="conv1 pharm h")
    close(from_cm(st.pharm_v), d["conv1_pharm_v"], what="conv1 pharm v")
    close(st.prot_h, d["conv1_prot_h"], what="conv1 prot h")
    close(from_cm(st.prot_v), d["conv1_prot_v"], what="conv1 prot v")
    close(eps_h, d["eps_h"], what="eps_h")
    close(eps_x, d["eps_x"], what="eps_x")


def test_denoiser_config1_size_against_oracle(env):
    """configs[0] shape: one 400-atom pocket x 30 samples of sizes 3..8."""
    from pharmacoforge_b200.synthetic import readme_sizes
    g, b = env.build([(400, 0)], [readme_sizes(30)])
    x, h, prot = random_state(b, 99, 4.0)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    tt = torch.full((30,), 0.42)
    wh, wx = env.O.denoiser(env.sd, b, tt, env.cfg)
    gh, gx = env.model.dynamics(g, tt, None)
    g.check_status()
    close(gh, wh, what="eps_h")
    close(gx, wx, what="eps_x")


def test_denoiser_large_pocket_config3_shape(env):
    """configs[2] shape (large-pocket stress): 1,500-atom pockets, pharmacophore sizes up to 16, per-graph timesteps."""
    g, b = env.build([(1500, 31), (1500, 32)], [[16, 3], [11]])
    x, h, prot = random_state(b, 7, 6.0)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    tt = torch.tensor([0.9, 0.13, 0.5])
    wh, wx = env.O.denoiser(env.sd, b, tt, env.cfg)
    gh, gx = env.model.dynamics(g, tt, None)
    g.check_status()
    close(gh, wh, what="eps_h")
    close(gx, wx, what="eps_x")


# ------------------------------------------------------------------------------------------------ K5b + loop
def test_posterior_step_bit_exact(env):
    O = env.O
    g, b = env.build([(100, 5)], [[4, 7, 3]])
    x, h, prot = random_state(b, 31)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x.clone(), h.clone(), prot.clone()
    gen = torch.Generator().manual_seed(4)
    nf = g.n_pharm
    eps_x, eps_h = torch.randn(nf, 3, generator=gen), torch.randn(nf, 6, generator=gen)
    nx, nh = torch.randn(nf, 3, generator=gen), torch.randn(nf, 6, generator=gen)
    t_host, a_ts, v_t, s_q = env.model.step_tables()[:4]
    for i in (0, 50, 99):
        a, v, q = (torch.tensor(float(val[i])) for val in (a_ts, v_t, s_q))
        fb = b.pharm_b
        b.pharm_x = (b.pharm_x / a - v * eps_x) + q * nx
        b.pharm_h = (b.pharm_h / a - v * eps_h) + q * nh
        O.remove_pharm_com(b)
        env.ops.posterior_step(g.pharm_x, g.pharm_h, eps_x.cuda(), eps_h.cuda(), nx.cuda(), nh.cuda(), g.pharm_ptr,
                               g.prot_x, g.prot_ptr, float(a_ts[i]), float(v_t[i]), float(s_q[i]))
        assert torch.equal(g.pharm_x.cpu(), b.pharm_x) and torch.equal(g.pharm_h.cpu(), b.pharm_h)
        assert torch.equal(g.prot_x.cpu(), b.prot_x)


def test_step_tables_match_reference_constants(env, golden):
    c = golden("constants.npz")
    t_host, a_ts, v_t, s_q = env.model.step_tables()[:4]
    assert np.array_equal(a_ts[::-1], c["alpha_ts"]) and np.array_equal(v_t[::-1], c["var_terms"])
    assert np.array_equal(s_q[::-1], c["sigma_q"])
    assert np.array_equal(env.model.gamma.gamma.detach().cpu().numpy(), c["gamma"])


def test_teacher_forced_steps_against_reference_trajectory(env, golden):
    """Feed the reference's state at step i, take ONE CUDA step with the reference's noise, compare with the
    reference's state at step i+1 (SURVEY.md §7 hard part 3: the gate that chaos cannot blur)."""
    d = golden("sample_traj.npz")
    sizes = [int(v) for v in d["sizes"]]
    n_atoms = int(d["n_atoms"])
    g, b = env.build([(n_atoms, int(d["pocket_seed"]))], [sizes])
    base = g.prot_x0.cpu()
    noise = t(d["noise"])
    model = env.model
    st = model.dynamics.bind(g)
    nx, nh = noise[:, :, 0:3].contiguous().cuda(), noise[:, :, 3:9].contiguous().cuda()
    for i in (0, 1, 2, 10, 37, 50, 73, 98, 99):
        # protein frame of graph j at step i = input pocket shifted so that its first atom matches the fixture
        prot = base.clone()
        for j in range(len(sizes)):
            sl = slice(j * n_atoms, (j + 1) * n_atoms)
            prot[sl] += t(d["traj_prot0"][i, j]) - prot[sl][0]
        g.prot_x.copy_(prot.cuda())
        g.pharm_x.copy_(t(d["traj_x"][i]).cuda())
        g.pharm_h.copy_(t(d["traj_h"][i]).cuda())
        model._run_steps(g, st, nx, nh, i, 1)
        g.check_status()
        close(g.pharm_x, d["traj_x"][i + 1], rtol=1e-4, atol=2e-5, what=f"x step {i}")
        close(g.pharm_h, d["traj_h"][i + 1], rtol=1e-4, atol=2e-5, what=f"h step {i}")


def test_full_reverse_diffusion_against_reference(env, golden):
    d = golden("sample_traj.npz")
    sizes = [int(v) for v in d["sizes"]]
    g, b = env.build([(int(d["n_atoms"]), int(d["pocket_seed"]))], [sizes])
    out = env.model.sample_given_receptor(g, noise=t(d["noise"]), visualize_trajectory=True)
    x = torch.cat([p.ph_coords for p in out])
    h = torch.cat([p.ph_feats for p in out])
    types = torch.cat([p.ph_feats_idxs for p in out])
    # 100 chained steps: fp32 re-association noise compounds, so the end-to-end bar is looser than per step
    close(x, d["final_x"], rtol=1e-3, atol=2e-3, what="final x")
    close(h, d["final_h"], rtol=1e-3, atol=2e-3, what="final h")
    assert np.array_equal(types.numpy(), d["final_type"])
    close(g.prot_x, d["final_prot"], rtol=1e-4, atol=1e-3, what="final prot frame")
    assert out[0].pos_frames.shape == (101, sizes[0], 3)
    assert out[0].to_xyz_file().splitlines()[0] == str(sizes[0])


def _sample_multi_inputs(env, d):
    pockets = [env.Pocket.from_numpy(*env.make_pocket(int(n), seed=int(s))) for n, s in d["pockets"]]
    flat, n_pharms, i = d["n_pharms_flat"].tolist(), [], 0
    for k in d["n_pharms_per_pocket"].tolist():
        n_pharms.append(flat[i:i + k])
        i += k
    return pockets, n_pharms


def test_sample_chunking_coms_and_frame_values_against_reference(env, golden):
    """`PharmacophoreDiff.sample` (pharmacodiff.py:516-578) against the fixture written by the reference's own `sample`:
    six graphs over three pockets, max_batch_size 4 (chunks of 4 + 2, the second one starting mid-pocket), an explicit
    per-pocket init_pharm_com indexed per graph, regrouping per pocket, and the VALUES of the trajectory frames
    (get_pos_feat_for_visual, :360-378: x_t shifted by init_prot_com - prot_com, h_t as is)."""
    d = golden("sample_multi.npz")
    pockets, n_pharms = _sample_multi_inputs(env, d)
    out = env.model.sample(pockets, n_pharms, max_batch_size=int(d["max_batch_size"]),
                           init_pharm_com=t(d["init_pharm_com"]), visualize_trajectory=True, noise=t(d["noise"]))
    assert [len(o) for o in out] == d["n_pharms_per_pocket"].tolist()
    ph = [p for o in out for p in o]
    assert [p.n_ph_centers for p in ph] == d["n_pharms_flat"].tolist()
    pos = torch.cat([p.pos_frames for p in ph], dim=1)
    feat = torch.cat([p.feat_frames for p in ph], dim=1)
    assert pos.shape == d["pos_frames"].shape and feat.shape == d["feat_frames"].shape
    # frame 0 is the injected z_T moved to the input frame; the first frames are a few steps deep: the per-step bar
    close(pos[:4], d["pos_frames"][:4], rtol=1e-4, atol=1e-4, what="first frames (pos)")
    close(feat[:4], d["feat_frames"][:4], rtol=1e-4, atol=2e-5, what="first frames (feat)")
    # 100 chained steps: the end-to-end bar of test_full_reverse_diffusion_against_reference
    close(pos, d["pos_frames"], rtol=1e-3, atol=2e-3, what="all frames (pos)")
    close(feat, d["feat_frames"], rtol=1e-3, atol=2e-3, what="all frames (feat)")
    close(torch.cat([p.ph_coords for p in ph]), d["final_x"], rtol=1e-3, atol=2e-3, what="final x")
    close(torch.cat([p.ph_feats for p in ph]), d["final_h"], rtol=1e-3, atol=2e-3, what="final h")
    assert np.array_equal(torch.cat([p.ph_feats_idxs for p in ph]).numpy(), d["final_type"])
    assert ph[0].to_xyz_file() == str(d["xyz_first"])            # writers: same text as the reference's
    assert ph[0].traj_to_xyz().splitlines()[:5] == str(d["traj_xyz_first"]).splitlines()[:5]


def test_sample_default_com_and_chunk_invariance(env):
    """sample() without init_pharm_com uses each pocket's mean position (pharmacodiff.py:531-535); the result of a graph
    does not depend on max_batch_size (bit-identical), checked against the oracle's sample_multi on a short run."""
    specs = [(120, 41), (90, 42)]
    n_pharms = [[3, 6, 4], [5, 8]]
    pk = [env.make_pocket(n, seed=s) for n, s in specs]
    pockets = [env.Pocket.from_numpy(p, h) for p, h in pk]
    nf = sum(sum(s) for s in n_pharms)
    noise = torch.randn(101, nf, 9, generator=torch.Generator().manual_seed(9))
    res = {}
    for mb in (2, 5, 32):
        out = env.model.sample(pockets, n_pharms, max_batch_size=mb, noise=noise, n_steps=6)
        res[mb] = (torch.cat([p.ph_coords for o in out for p in o]), torch.cat([p.ph_feats for o in out for p in o]))
    for mb in (5, 32):
        assert torch.equal(res[mb][0], res[2][0]) and torch.equal(res[mb][1], res[2][1]), mb
    want = env.O.sample_multi(env.sd, [(t(p), t(h)) for p, h in pk], n_pharms, noise, 100, env.sd["gamma.gamma"],
                              env.cfg, max_batch_size=2, steps=6)
    close(res[2][0], torch.cat([p["x"] for o in want for p in o]), rtol=1e-4, atol=1e-4, what="x after 6 steps")
    close(res[2][1], torch.cat([p["h"] for o in want for p in o]), rtol=1e-4, atol=2e-5, what="h after 6 steps")


def test_configs2_shape_full_reverse_diffusion(env):
    """BASELINE.json configs[2] shape through all 100 steps: a 1,500-atom pocket, pharmacophores of 16 and 3 centres.
    Teacher-forced from the oracle's own trajectory at several steps (per-step bar 1e-4), then end to end."""
    O, model = env.O, env.model
    specs, sizes = [(1500, 61)], [[16, 3]]
    g, b = env.build(specs, sizes)
    noise = torch.randn(101, 19, 9, generator=torch.Generator().manual_seed(5))
    rec = []
    wx, wh, wt, wprot = O.sample(env.sd, b, noise, 100, env.sd["gamma.gamma"], env.cfg, record=rec)
    st = model.dynamics.bind(g)
    nx, nh = noise[:, :, 0:3].contiguous().cuda(), noise[:, :, 3:9].contiguous().cuda()
    for i in (0, 25, 60, 99):
        g.pharm_x.copy_(rec[i][0].cuda())
        g.pharm_h.copy_(rec[i][1].cuda())
        g.prot_x.copy_(rec[i][2].cuda())
        model._run_steps(g, st, nx, nh, i, 1)
        g.check_status()
        close(g.pharm_x, rec[i + 1][0], rtol=1e-4, atol=2e-5, what=f"x step {i}")
        close(g.pharm_h, rec[i + 1][1], rtol=1e-4, atol=2e-5, what=f"h step {i}")
    g.prot_x.copy_(g.prot_x0)
    x, h = model.sample_given_receptor(g, noise=noise, return_tensors=True)
    close(x, wx, rtol=1e-3, atol=2e-3, what="final x")
    close(h, wh, rtol=1e-3, atol=2e-3, what="final h")
    assert torch.equal(h.argmax(dim=1).cpu(), wt)


def test_bench_scale_batch_spot_check(env):
    """The batch bench.py times (BASELINE.json configs[1]: 256 x 400-atom pockets x 30 samples = 7,680 graphs, 3.07 M
    protein nodes, ~23 M pp edges): one teacher-forced reverse step at that scale, three graphs drawn at random compared
    with the oracle run on each graph alone (graphs are independent, so the oracle needs only those three), and
    bit-identity of the same three graphs against a small batch that contains only them."""
    from pharmacoforge_b200.synthetic import readme_sizes
    O, model = env.O, env.model
    n_pockets, atoms = 256, 400
    pk = [env.make_pocket(atoms, seed=i) for i in range(n_pockets)]
    sizes = [readme_sizes(30)] * n_pockets
    g = env.GraphBatch.from_pockets([env.Pocket.from_numpy(p, h) for p, h in pk], sizes, env.dev)
    assert g.n_graphs == 7680 and g.n_prot == 7680 * atoms and g.n_pp_edges > 20_000_000
    gen = torch.Generator().manual_seed(2024)
    x = torch.randn(g.n_pharm, 3, generator=gen) * 3.0
    h = torch.randn(g.n_pharm, 6, generator=gen)
    noise = torch.randn(2, g.n_pharm, 9, generator=gen)
    fptr, pptr = g.pharm_ptr_host.astype(np.int64), g.prot_ptr_host.astype(np.int64)
    # sampler frame: every graph's protein centred on its pharmacophore + a small offset
    com = torch.from_numpy(np.stack([p.mean(axis=0) for p, _ in pk])).float()
    off = torch.randn(g.n_graphs, 3, generator=gen)
    gi_of_prot = torch.repeat_interleave(torch.arange(g.n_graphs), atoms)
    prot = g.prot_x0.cpu() - com[gi_of_prot // 30] + off[gi_of_prot]
    st = model.dynamics.bind(g)
    step = 57
    g.pharm_x.copy_(x.cuda())
    g.pharm_h.copy_(h.cuda())
    g.prot_x.copy_(prot.cuda())
    # _run_steps(first=step) reads noise row 1 + step: place the drawn row there
    nx = torch.zeros(step + 2, g.n_pharm, 3, device=env.dev)
    nh = torch.zeros(step + 2, g.n_pharm, 6, device=env.dev)
    nx[step + 1], nh[step + 1] = noise[1, :, 0:3].cuda(), noise[1, :, 3:9].cuda()
    model._run_steps(g, st, nx, nh, step, 1)
    g.check_status()
    gx, gh, gprot = g.pharm_x.cpu(), g.pharm_h.cpu(), g.prot_x.cpu()
    for gidx in (3, 4097, 7679):
        p = gidx // 30
        fs, ps = slice(int(fptr[gidx]), int(fptr[gidx + 1])), slice(int(pptr[gidx]), int(pptr[gidx + 1]))
        b = O.build_batch([(t(pk[p][0]), t(pk[p][1]))], [[sizes[p][gidx % 30]]])
        b.prot_x, b.pharm_x, b.pharm_h = prot[ps].clone(), x[fs].clone(), h[fs].clone()
        O.reverse_step(env.sd, b, 99 - step, 100, env.sd["gamma.gamma"], env.cfg, noise[1, fs, 0:3], noise[1, fs, 3:9])
        close(gx[fs], b.pharm_x, rtol=1e-4, atol=2e-5, what=f"graph {gidx} x")
        close(gh[fs], b.pharm_h, rtol=1e-4, atol=2e-5, what=f"graph {gidx} h")
        close(gprot[ps], b.prot_x, rtol=1e-4, atol=1e-4, what=f"graph {gidx} prot frame")
        # the same graph alone in a batch of one: bit-identical to its rows in the 7,680-graph batch
        g1 = env.GraphBatch.from_pockets([env.Pocket.from_numpy(*pk[p])], [[sizes[p][gidx % 30]]], env.dev)
        st1 = model.dynamics.bind(g1)
        g1.pharm_x.copy_(x[fs].cuda())
        g1.pharm_h.copy_(h[fs].cuda())
        g1.prot_x.copy_(prot[ps].cuda())
        model._run_steps(g1, st1, nx[:, fs].contiguous(), nh[:, fs].contiguous(), step, 1)
        assert torch.equal(g1.pharm_x.cpu(), gx[fs]) and torch.equal(g1.pharm_h.cpu(), gh[fs]), gidx
    del g, st
    torch.cuda.empty_cache()


def test_philox_noise_and_cuda_graph_loop(env):
    """Throughput path (noise=None): in-kernel Philox noise keyed from torch's CUDA generator, and the whole T-step loop
    replayed as one CUDA graph.  Draws are standard normal and independent across streams / steps; a run is reproducible
    under torch.manual_seed; the graph replay is bit-identical to the eager enqueue; and one posterior step with Philox
    noise equals the bit-exact injected-noise step fed with the same draws."""
    ops, model = env.ops, env.model
    seed = torch.tensor([123456789], dtype=torch.int64, device=env.dev)
    n = 1 << 20
    draws = {}
    for key in ((0, 0), (1, 0), (0, 7)):
        out = torch.empty(n, device=env.dev)
        ops.philox_normal(out, seed, key[0], key[1])
        draws[key] = out
        assert abs(float(out.mean())) < 5e-3 and abs(float(out.var()) - 1.0) < 5e-3
        assert abs(float((out ** 4).mean()) - 3.0) < 5e-2 and float(out.abs().max()) < 6.5
    for a, b in (((0, 0), (1, 0)), ((0, 0), (0, 7))):
        assert abs(float((draws[a] * draws[b]).mean())) < 5e-3          # streams / steps are uncorrelated
    assert abs(float((draws[(0, 0)][1:] * draws[(0, 0)][:-1]).mean())) < 5e-3
    out2 = torch.empty(1000, device=env.dev)
    ops.philox_normal(out2, seed, 0, 0)
    assert torch.equal(out2, draws[(0, 0)][:1000])                       # a pure function of (seed, stream, step, index)

    g, b = env.build([(150, 3), (90, 4)], [[3, 5, 8], [6, 4]])
    st = model.dynamics.bind(g)
    # one Philox step == the injected-noise step with the same draws (the bit-exact posterior kernel)
    x, h, prot = random_state(b, 13)
    eps_x, eps_h = torch.randn(g.n_pharm, 3, device=env.dev), torch.randn(g.n_pharm, 6, device=env.dev)
    nz_x, nz_h = torch.empty(g.n_pharm, 3, device=env.dev), torch.empty(g.n_pharm, 6, device=env.dev)
    ops.philox_normal(nz_x, seed, 0, 5)
    ops.philox_normal(nz_h, seed, 1, 5)
    res = []
    for philox in (False, True):
        px, ph, pp = x.cuda().clone(), h.cuda().clone(), prot.cuda().clone()
        if philox:
            import ctypes as C
            from pharmacoforge_b200 import _lib
            L = _lib.load()
            vp = lambda tt: C.c_void_p(tt.data_ptr())
            _lib.check(L.pf_posterior_step_philox(vp(px), vp(ph), 6, vp(eps_x), vp(eps_h), vp(seed), 5, vp(g.pharm_ptr), vp(pp),
                                                  vp(g.prot_ptr), g.n_graphs, 0.97, 0.11, 0.23,
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "philox step")
        else:
            ops.posterior_step(px, ph, eps_x, eps_h, nz_x, nz_h, g.pharm_ptr, pp, g.prot_ptr, 0.97, 0.11, 0.23)
        res.append((px, ph, pp))
    assert all(torch.equal(u, v) for u, v in zip(res[0], res[1]))

    def run(graph):
        model.use_cuda_graph = graph
        torch.manual_seed(77)
        g.prot_x.copy_(g.prot_x0)
        xx, hh = model.sample_given_receptor(g, n_steps=12, return_tensors=True)
        return xx.clone(), hh.clone()
    try:
        e1, e2 = run(False), run(False)
        assert torch.equal(e1[0], e2[0]) and torch.equal(e1[1], e2[1])          # reproducible under manual_seed
        g1 = run(True)                                                           # captures (the eager runs warmed up)
        g2 = run(True)                                                           # replays
        assert len(st.graphs) == 1
        for r in (g1, g2):
            assert torch.equal(r[0], e1[0]) and torch.equal(r[1], e1[1])
        torch.manual_seed(78)
        g.prot_x.copy_(g.prot_x0)
        other = model.sample_given_receptor(g, n_steps=12, return_tensors=True)[0]
        assert not torch.equal(other, e1[0])                                     # a new key gives a new trajectory
        assert torch.isfinite(other).all()
    finally:
        model.use_cuda_graph = False


# ------------------------------------------------------------------------------------------------ properties
def test_determinism_and_batch_composition_invariance(env):
    from pharmacoforge_b200.synthetic import readme_sizes
    model = env.model
    specs = [(400, 0), (250, 1), (120, 2)]
    sizes = [readme_sizes(6), [4, 8], [3, 5, 7]]
    g, b = env.build(specs, sizes)
    noise = torch.randn(11, g.n_pharm, 9, generator=torch.Generator().manual_seed(1))
    x1, h1 = model.sample_given_receptor(g, noise=noise, n_steps=10, return_tensors=True)
    g2, _ = env.build(specs, sizes)
    x2, h2 = model.sample_given_receptor(g2, noise=noise, n_steps=10, return_tensors=True)
    assert torch.equal(x1, x2) and torch.equal(h1, h2), "not bit-reproducible"
    # the middle pocket alone must give bit-identical results: a graph never sees its batch neighbours
    lo, hi = int(g.pharm_ptr_host[6]), int(g.pharm_ptr_host[8])
    g3, _ = env.build([specs[1]], [sizes[1]])
    x3, h3 = model.sample_given_receptor(g3, noise=noise[:, lo:hi].contiguous(), n_steps=10, return_tensors=True)
    assert torch.equal(x3, x1[lo:hi]) and torch.equal(h3, h1[lo:hi])


def test_shard_count_invariance(env):
    """Results for N graphs as 1xN equal 2x(N/2) and 3 uneven shards bit for bit (multi-GPU sharding contract)."""
    from pharmacoforge_b200.sharding import shard_ranges
    model = env.model
    pk = [env.make_pocket(n, seed=s) for n, s in [(200, 0), (300, 1)]]
    pockets = [env.Pocket.from_numpy(p, h) for p, h in pk]
    sizes = [[3, 4, 5, 6], [7, 8, 3, 4, 5]]
    full = env.GraphBatch.from_pockets(pockets, sizes, env.dev)
    noise = torch.randn(6, full.n_pharm, 9, generator=torch.Generator().manual_seed(2))
    xf, hf = model.sample_given_receptor(full, noise=noise, n_steps=5, return_tensors=True)
    ptr = full.pharm_ptr_host
    for world in (2, 3):
        xs = []
        for r in range(world):
            rng = shard_ranges(sizes, world)[r]
            gb = env.GraphBatch.from_pockets(pockets, sizes, env.dev, graph_range=rng)
            n = noise[:, int(ptr[rng.start]):int(ptr[rng.stop])].contiguous()
            xs.append(model.sample_given_receptor(gb, noise=n, n_steps=5, return_tensors=True)[0])
        assert torch.equal(torch.cat(xs), xf), world


def test_equivariance_of_eps(env):
    """SE(3): rotating + translating every graph rotates eps_x and leaves eps_h unchanged (up to fp32 noise)."""
    g, b = env.build([(200, 3)], [[5, 8]])
    x, h, prot = random_state(b, 77)
    tt = torch.tensor([0.3, 0.8])
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    eh, ex = (v.clone() for v in env.model.dynamics(g, tt, None))
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(5)))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    shift = torch.tensor([1.5, -2.0, 0.7])
    env.set_state(g, b, (x @ q.T + shift).cuda(), h.cuda(), (prot @ q.T + shift).cuda())
    eh2, ex2 = env.model.dynamics(g, tt, None)
    close(eh2, eh, rtol=1e-3, atol=1e-4, what="eps_h invariance")
    close(ex2, ex.cpu() @ q.T, rtol=1e-3, atol=1e-4, what="eps_x equivariance")


def test_degree_overflow_is_reported(env):
    pos = np.random.default_rng(0).uniform(0, 1.0, size=(80, 3)).astype(np.float32)   # 79 neighbours each
    onehot = np.zeros((80, 11), dtype=np.float32)
    onehot[:, 0] = 1
    from pharmacoforge_b200._lib import PfError
    with pytest.raises(PfError):     # 79 in-edges do not fit a 64-row tile of the FFMA kernels: reported at batch build
        env.GraphBatch.from_pockets([env.Pocket.from_numpy(pos, onehot)], [[3]], env.dev, tile_rows=64)
    g = env.GraphBatch.from_pockets([env.Pocket.from_numpy(pos, onehot)], [[3]], env.dev, tile_rows=128)
    g.check_status()                 # max_num_neighbors=100 always fits the 128-row tcgen05 tile


def test_forward_loss_matches_reference(env, golden, sd, dyn_cfg):
    """PharmacophoreDiff.forward (the training / validation objective, pharmacodiff.py:162-243) on the CUDA denoiser
    against the reference's own forward (golden fixture) and the oracle, with the same injected (t, eps)."""
    import pf_oracle as O
    from pharmacoforge_b200.synthetic import make_pocket
    g = golden("forward_loss.npz")
    sizes = list(map(int, g["sizes"]))
    pos, onehot = make_pocket(int(g["n_atoms"]), seed=int(g["pocket_seed"]))
    gb = env.GraphBatch.from_pockets([env.Pocket.from_numpy(pos, onehot)], [sizes], env.dev)
    gb.set_pharmacophores(t(g["x0"]), t(g["h0"]))
    losses, metrics = env.model.validation_step(gb, t_int=t(g["t_int"]), eps={"x": t(g["eps_x"]), "h": t(g["eps_h"])})
    b = O.build_batch([(t(pos), t(onehot))], [sizes])
    lo, mo = O.forward_loss(sd, b, t(g["x0"]), t(g["h0"]), t(g["t_int"]), t(g["eps_x"]), t(g["eps_h"]), 100,
                            sd["gamma.gamma"], dyn_cfg, phase="val")
    for k, v in {**lo, **mo}.items():
        got = float({**losses, **metrics}[k])
        ref = float(g[k.replace(" ", "_")])
        tol = 2e-4 * max(1.0, abs(ref))   # sums of squares of eps errors that are each within 1e-4
        assert abs(got - ref) <= tol and abs(got - float(v)) <= tol, (k, got, ref, float(v))
    assert abs(float(losses["val total loss"]) - float(g["val_pos_loss"]) - float(g["val_feat_loss"])) < 1e-3


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_shared_pocket_messages_match_nominal(env, precision):
    """dynamics.share_pocket_messages (SURVEY.md hard part 5a + 5b + 5c, opt-in): first-layer pp messages once per distinct
    pocket, protein rows encoded / updated only where the last layer's pf edges read them.  eps equals the nominal path up
    to fp32 rounding of x_src - x_dst (bar 2e-5, well inside the 1e-4 parity bar), over a ragged batch with several samples
    per pocket, a graph_range that starts mid-pocket, and a short sampling run; per-graph timesteps are rejected."""
    dyn = env.model.dynamics
    pk = [env.make_pocket(n, seed=s) for n, s in ((400, 0), (120, 2), (250, 1))]
    pockets = [env.Pocket.from_numpy(p, h) for p, h in pk]
    sizes = [[3, 8, 5, 4], [4, 6], [7, 3, 5]]
    gr = range(1, 9)
    g = env.GraphBatch.from_pockets(pockets, sizes, env.dev, graph_range=gr)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(g.n_pharm, 3, generator=gen) * 3.0
    h = torch.randn(g.n_pharm, 6, generator=gen)
    # sampler frame: proteins centred near their pharmacophores
    bidx = g.batch_idxs()["prot"].cpu()
    com = torch.stack([g.prot_x0.cpu()[bidx == i].mean(0) for i in range(g.n_graphs)])
    prot = g.prot_x0.cpu() - com[bidx] + torch.randn(g.n_graphs, 3, generator=gen)[bidx]
    tt = torch.full((g.n_graphs,), 0.37)
    old = (dyn.edge_mlp_precision, dyn.share_pocket_messages, dyn.skip_dead_work)
    try:
        dyn.edge_mlp_precision = precision
        res = {}
        for share in (False, True):
            dyn.share_pocket_messages = share
            g.pharm_h = h.cuda().clone()
            g.pharm_x.copy_(x.cuda())
            g.prot_x.copy_(prot.cuda())
            eh, ex = dyn(g, tt, None)
            res[share] = (eh.clone(), ex.clone())
        st = dyn.bind(g)
        assert st.share is not None and (st.args.flags & 8)
        if precision == "fp32":
            close(res[True][0], res[False][0], rtol=2e-5, atol=2e-6, what="eps_h shared vs nominal")
            close(res[True][1], res[False][1], rtol=2e-5, atol=2e-6, what="eps_x shared vs nominal")
        else:   # single-pass fp16: rounding of x_diff moves fp16 operands by an ulp here and there
            within(res[True][0], res[False][0], 2e-3, what="eps_h shared vs nominal (fp16)")
            within(res[True][1], res[False][1], 2e-3, what="eps_x shared vs nominal (fp16)")
        with pytest.raises(ValueError):
            dyn(g, torch.linspace(0.1, 0.9, g.n_graphs), None)
        # a short reverse-diffusion run in both modes with the same injected noise
        noise = torch.randn(9, g.n_pharm, 9, generator=gen)
        runs = {}
        for share in (False, True):
            dyn.share_pocket_messages = share
            g.prot_x.copy_(g.prot_x0)
            runs[share] = env.model.sample_given_receptor(g, noise=noise, n_steps=8, return_tensors=True)
        tol = dict(rtol=1e-4, atol=1e-4) if precision == "fp32" else dict(rtol=5e-2, atol=5e-2)
        close(runs[True][0], runs[False][0], what="x after 8 steps", **tol)
        close(runs[True][1], runs[False][1], what="h after 8 steps", **tol)
    finally:
        dyn.edge_mlp_precision, dyn.share_pocket_messages, dyn.skip_dead_work = old


def test_dead_work_elimination_is_bit_exact(env):
    """skip_dead_work drops the last layer's protein-side kernels (never read, dynamics_gvp.py:84-92): the sampled
    pharmacophores must not change by a single bit."""
    g, _ = env.build([(150, 21), (90, 22)], [[3, 6], [8, 4]])
    noise = torch.randn(7, g.n_pharm, 9, generator=torch.Generator().manual_seed(5))
    x_a, h_a = env.model.sample_given_receptor(g, noise=noise, n_steps=6, return_tensors=True)
    g2, _ = env.build([(150, 21), (90, 22)], [[3, 6], [8, 4]])
    env.model.dynamics.skip_dead_work = True
    try:
        x_b, h_b = env.model.sample_given_receptor(g2, noise=noise, n_steps=6, return_tensors=True)
    finally:
        env.model.dynamics.skip_dead_work = False
    assert torch.equal(x_a, x_b) and torch.equal(h_a, h_b)


# ------------------------------------------------------------------------------------------------ fp16 single-pass mode
@pytest.mark.parametrize("etype_idx", [0, 1, 2, 3])
@pytest.mark.parametrize("with_vectors", [False, True])
def test_fp16_single_pass_edge_conv(env, etype_idx, with_vectors):
    """pf_edge_conv_tc_f16 against the fp32 oracle at the stated reduced-precision tolerance."""
    _edge_conv_case(env, etype_idx, 1 if with_vectors else 0, with_vectors, "tc", fp16=True)


def test_fp16_single_pass_node_update(env):
    O, ops = env.O, env.ops
    gen = torch.Generator().manual_seed(18)
    for n in (1, 129, 5000):
        h, v = torch.randn(n, 128, generator=gen), torch.randn(n, 16, 3, generator=gen)
        ah, av = torch.randn(n, 128, generator=gen), torch.randn(n, 16, 3, generator=gen)
        p = "dynamics.noise_predictor.conv_layers.1"
        s, vv = O.gvp_layernorm(env.sd, f"{p}.message_layer_norms.prot", h + ah, v + av)
        rs, rv = s, vv
        for i in range(2):
            rs, rv = O.gvp(env.sd, f"{p}.node_update_fns.prot.{i}", rs, rv)
        want_h, want_v = O.gvp_layernorm(env.sd, f"{p}.update_layer_norms.prot", s + rs, vv + rv)
        hd, vd = h.cuda(), to_cm(v).cuda()
        ops.node_update_tc(hd, vd, ah.cuda(), to_cm(av).cuda(), env.W.tcu_view(1, 1), hd, vd, True)
        torch.cuda.synchronize()
        within(hd, want_h, FP16_TOL, what=f"fp16 node_update h n={n}")
        within(from_cm(vd), want_v, FP16_TOL, what=f"fp16 node_update v n={n}")


def test_fp16_single_pass_denoiser_and_sampling(env, golden):
    """The whole denoiser in the reduced-precision mode: eps within FP16_TOL of the oracle's fp32 eps, edge lists
    unchanged (the graph kernels stay fp32), and a full 100-step reverse diffusion that stays close to the fp32
    trajectory of the reference fixture.  The mode must be opt-in and must actually change the arithmetic."""
    from pharmacoforge_b200.synthetic import readme_sizes
    dyn = env.model.dynamics
    g, b = env.build([(400, 0)], [readme_sizes(30)])
    x, h, prot = random_state(b, 99, 4.0)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    tt = torch.full((30,), 0.42)
    wh, wx = env.O.denoiser(env.sd, b, tt, env.cfg)
    assert dyn.edge_mlp_precision == "fp32"
    fh, fx = (v.clone() for v in dyn(g, tt, None))
    e32 = g.dynamic_edges()
    try:
        dyn.edge_mlp_precision = "fp16"
        gh, gx = dyn(g, tt, None)
        g.check_status()
        rh = within(gh, wh, FP16_TOL, what="fp16 eps_h")
        rx = within(gx, wx, FP16_TOL, what="fp16 eps_x")
        assert not torch.equal(gh, fh), "the fp16 flag did not reach the kernels"
        assert rh > 1e-6 or rx > 1e-6
        e16 = g.dynamic_edges()
        for et in ("ff", "pf", "fp"):
            assert all(np.array_equal(a, c) for a, c in zip(canon(*e32[et]), canon(*e16[et]))), et
        d = golden("sample_traj.npz")
        sizes = [int(v) for v in d["sizes"]]
        g2, _ = env.build([(int(d["n_atoms"]), int(d["pocket_seed"]))], [sizes])
        x16, h16 = env.model.sample_given_receptor(g2, noise=t(d["noise"]), return_tensors=True)
        # 100 chained steps at ~1e-3 per call: the samples stay within a fraction of an Angstrom of the fp32 ones
        dx = (x16.cpu() - t(d["final_x"])).abs().max().item()
        dh = (h16.cpu() - t(d["final_h"])).abs().max().item()
        assert dx < 0.25 and dh < 0.25, (dx, dh)
    finally:
        dyn.edge_mlp_precision = "fp32"


def test_edge_cases_max_size_graph_and_empty_size_lists(env):
    """Limits of the batch layout: a graph with PF_MAX_PHARM_PER_GRAPH = 128 centres (ff in-degree 127, one full tile) next to a
    one-centre graph through the fused denoiser against the oracle; 129 centres are refused on the host; a pocket with an empty
    size list contributes no graph and `sample` regroups around it (pharmacodiff.py:538-576)."""
    g, b = env.build([(400, 0), (40, 3)], [[128, 1], [2]])
    x, h, prot = random_state(b, 31, 2.5)
    env.set_state(g, b, x.cuda(), h.cuda(), prot.cuda())
    b.pharm_x, b.pharm_h, b.prot_x = x, h, prot
    tt = torch.tensor([0.4, 0.4, 0.9])
    wh, wx = env.O.denoiser(env.sd, b, tt, env.cfg)
    gh, gx = env.model.dynamics(g, tt, None)
    close(gh, wh, what="eps_h (128 centres)")
    close(gx, wx, what="eps_x (128 centres)")
    assert int(g.ff_cnt.max()) == 127
    with pytest.raises(ValueError):
        env.build([(400, 0)], [[129]])
    pockets = [env.Pocket.from_numpy(*env.make_pocket(n, seed=s)) for n, s in ((120, 1), (90, 2), (60, 3))]
    sizes = [[3, 4], [], [5]]
    noise = torch.randn(6, 12, 9, generator=torch.Generator().manual_seed(2))
    out = env.model.sample(pockets, sizes, max_batch_size=2, noise=noise, n_steps=5)
    assert [len(o) for o in out] == [2, 0, 1] and [p.n_ph_centers for o in out for p in o] == [3, 4, 5]
    # the same graphs without the empty pocket: identical samples
    out2 = env.model.sample([pockets[0], pockets[2]], [[3, 4], [5]], max_batch_size=2, noise=noise, n_steps=5)
    for a_, b_ in zip([p for o in out for p in o], [p for o in out2 for p in o]):
        assert torch.equal(a_.ph_coords, b_.ph_coords) and torch.equal(a_.ph_feats, b_.ph_feats)
