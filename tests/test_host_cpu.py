"""CPU-only checks of the host side: the C-ABI library loads and exports everything the header declares, the
Python mirror keeps the reference's state_dict layout, packing, sharding and the synthetic generator."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pharmacoforge_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pharmacoforge_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|size_t|const char\*)\s+(pf_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.pf_abi_version() == _lib.ABI_VERSION
    assert lib.pf_sample_args_size() == ctypes.sizeof(_lib.PfSampleArgs)
    assert lib.pf_scan_workspace_bytes(5000) >= 4 * 4


def test_gvp_layout_is_16_byte_aligned():
    from pharmacoforge_b200.weights import gvp_layout
    for dims in [(17, 16, 144, 128), (16, 16, 128, 128), (16, 1, 128, 64)]:
        offs, total = gvp_layout(*dims)
        assert all(o % 4 == 0 for o in offs) and total % 4 == 0
        vi, vo, si, so = dims
        vh = max(vi, vo)
        assert offs[1] - offs[0] >= vi * vh and offs[3] - offs[2] >= (si + vh) * so


def test_state_dict_layout_matches_reference(layout):
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    from conftest import DEV_DYNAMICS
    dyn = dict(DEV_DYNAMICS)
    cut = dyn.pop("graph_cutoffs")
    m = PharmacophoreDiff(6, 11, ["a", "b", "c", "d", "e", "f"], n_timesteps=100, graph_config={"graph_cutoffs": cut},
                          dynamics_config=dyn, precision=1e-5, rl_dist_threshold=0)
    sd = m.state_dict()
    assert set(sd) == set(layout)
    for k, shape in layout.items():
        assert list(sd[k].shape) == shape, k
    assert sum(v.numel() for v in sd.values()) == 772916
    assert not m.gamma.gamma.requires_grad


def test_checkpoint_round_trip(tmp_path, sd):
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    from conftest import DEV_DYNAMICS
    dyn = dict(DEV_DYNAMICS)
    cut = dyn.pop("graph_cutoffs")
    m = PharmacophoreDiff(6, 11, ["a", "b", "c", "d", "e", "f"], n_timesteps=100, graph_config={"graph_cutoffs": cut},
                          dynamics_config=dyn, precision=1e-5)
    m.load_state_dict(sd, strict=True)
    m.save_checkpoint(tmp_path / "last.ckpt")
    m2 = PharmacophoreDiff.load_from_checkpoint(tmp_path / "last.ckpt")
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k


def test_schedule_tables_match_reference_constants(golden, sd):
    from pharmacoforge_b200.diffusion import PharmacophoreDiff, polynomial_gamma
    c = golden("constants.npz")
    assert np.array_equal(polynomial_gamma(100, 1e-5, 2.0).numpy(), c["gamma"])
    from conftest import DEV_DYNAMICS
    dyn = dict(DEV_DYNAMICS)
    cut = dyn.pop("graph_cutoffs")
    m = PharmacophoreDiff(6, 11, ["a", "b", "c", "d", "e", "f"], n_timesteps=100, graph_config={"graph_cutoffs": cut},
                          dynamics_config=dyn, precision=1e-5)
    t_host, a_ts, v_t, s_q = m.step_tables()[:4]
    assert np.array_equal(a_ts[::-1], c["alpha_ts"]) and np.array_equal(v_t[::-1], c["var_terms"])
    assert np.array_equal(s_q[::-1], c["sigma_q"])
    assert t_host[0] == np.float32(1.0) and t_host[-1] == np.float32(0.01)


def test_weight_packing_places_transposed_blocks(sd):
    from pharmacoforge_b200.weights import gvp_layout, pack_gvp, pack_noise_head, pack_update
    p = "dynamics.noise_predictor.conv_layers.0.edge_message_fns.prot_pp_prot.0"
    w = pack_gvp(sd, p)
    offs, total = gvp_layout(17, 16, 144, 128)
    assert w.numel() == total
    Wf = sd[p + ".to_feats_out.0.weight"]
    assert torch.equal(w[offs[2]:offs[2] + 161 * 128].view(161, 128), Wf.t())
    assert torch.all(w[offs[2] + 161 * 128:offs[3]] == 0)      # K padding rows
    assert torch.equal(w[offs[4]:offs[4] + 128 * 16].view(128, 16), sd[p + ".scalar_to_vector_gates.weight"].t())
    assert pack_update(sd, "dynamics.noise_predictor.conv_layers.1", "prot", 2).numel() == 4 * 128 + 2 * gvp_layout(16, 16, 128, 128)[1]
    assert pack_noise_head(sd, "dynamics.noise_predictor.noise_predictor", 4).numel() % 4 == 0


def test_unsupported_configs_fail_loudly():
    from pharmacoforge_b200.dynamics import PharmRecDynamicsGVP
    with pytest.raises(NotImplementedError):
        PharmRecDynamicsGVP(6, 11, vector_size=8, n_convs=2, graph_cutoffs={"ff": 9}, message_norm="mean", pf_k=5)
    # a numeric message_norm (sum aggregation / norm; 0 = per-graph edges per node + 1) and pf_k = 0 (radius pf edges) are
    # built; dicts are not (they fail in the reference's own constructor check)
    assert PharmRecDynamicsGVP(6, 11, n_convs=2, graph_cutoffs={"ff": 9}, message_norm=10, pf_k=5).message_norm == 10
    assert PharmRecDynamicsGVP(6, 11, n_convs=2, graph_cutoffs={"ff": 9}, message_norm=0, pf_k=5).message_norm == 0
    with pytest.raises(ValueError):
        PharmRecDynamicsGVP(6, 11, n_convs=2, graph_cutoffs={"ff": 9}, message_norm=-1, pf_k=5)
    with pytest.raises(NotImplementedError):
        PharmRecDynamicsGVP(6, 11, n_convs=2, graph_cutoffs={"ff": 9}, message_norm={"prot": 1, "pharm": 1}, pf_k=5)
    assert PharmRecDynamicsGVP(6, 11, n_convs=2, graph_cutoffs={"ff": 9, "pf": 8}, message_norm="mean", pf_k=0).pf_k == 0
    with pytest.raises(KeyError):       # pf_k = 0 reads graph_cutoffs['pf'] (dynamics_gvp.py:211)
        PharmRecDynamicsGVP(6, 11, n_convs=2, graph_cutoffs={"ff": 9}, message_norm="mean", pf_k=0)
    with pytest.raises(ValueError):
        PharmRecDynamicsGVP(6, 11, n_convs=2, graph_cutoffs={"ff": 9}, message_norm="mean", pf_k=-1)


def test_ops_refuse_cpu_tensors():
    from pharmacoforge_b200 import _lib, ops
    with pytest.raises(_lib.PfError):
        ops.exclusive_scan(torch.zeros(4, dtype=torch.int32))


def test_shard_ranges_cover_and_balance():
    from pharmacoforge_b200.sharding import shard_ranges
    sizes = [[3] * 30 for _ in range(7)]
    for world in (1, 2, 3, 4, 8):
        rs = shard_ranges(sizes, world, pocket_atoms=[400] * 7)
        assert rs[0].start == 0 and rs[-1].stop == 210
        assert all(rs[i].stop == rs[i + 1].start for i in range(world - 1))
        assert max(len(r) for r in rs) - min(len(r) for r in rs) <= 1
    rs = shard_ranges([[3, 4], [5]], 8)
    assert sum(len(r) for r in rs) == 3
    assert shard_ranges([], 2) == [range(0, 0), range(0, 0)]


def test_synthetic_pocket_properties():
    from pharmacoforge_b200.synthetic import make_pocket, pocket_radius, readme_sizes
    pos, onehot = make_pocket(400, seed=0)
    assert pos.shape == (400, 3) and onehot.shape == (400, 11) and pos.dtype == np.float32
    assert np.array_equal(onehot.sum(1), np.ones(400))
    d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) + np.eye(400) * 10
    assert d.min() >= 1.3 - 1e-4
    assert abs(pocket_radius(400) - 12.2) < 0.1 and abs(pocket_radius(1500) - 18.7) < 0.1
    pos2, _ = make_pocket(400, seed=0)
    assert np.array_equal(pos, pos2)
    assert readme_sizes(30) == [3, 4, 5, 6, 7, 8] * 5


def test_tc_message_blob_layout(sd):
    """pack_message_tc: the UMMA K-slab images decode back to the reference weights (hi + lo), constants in place."""
    import torch
    from pharmacoforge_b200 import _lib, weights as W
    cp = "dynamics.noise_predictor.conv_layers.1"
    blob = W.pack_message_tc(sd, cp, "prot_pp_prot")
    assert blob.dtype == torch.uint8 and blob.numel() == _lib.load().pf_tc_msg_blob_bytes()

    def decode(img_bytes, n):      # inverse of umma_b_image for one [n, 16] K-slab
        return img_bytes.view(torch.float16).reshape(2, n // 8, 8, 8).permute(1, 2, 0, 3).reshape(n, 16).float()

    off = 0
    for g, nslab in enumerate(W.TC_SLABS):
        Wf = sd[f"{cp}.edge_message_fns.prot_pp_prot.{g}.to_feats_out.0.weight"].float()
        rec = torch.zeros(128, 16 * nslab)
        for s_ in range(nslab):
            hi = decode(blob[off:off + 4096], 128)
            lo = decode(blob[off + 4096:off + 8192], 128)
            rec[:, 16 * s_:16 * s_ + 16] = hi + lo
            off += 8192
        k = Wf.shape[1]
        # the images carry k * Wf, k = -log2(e): the kernels' SiLU reads t = k (Wf s + bf) from the accumulator
        assert torch.allclose(rec[:, :k], Wf * W.TC_PRESCALE, rtol=2 ** -20, atol=1e-7)
        assert float(rec[:, k:].abs().max() if k < rec.shape[1] else 0) == 0
    assert off == W.TC_SMALL_OFF
    consts = blob[W.TC_SMALL_OFF + W.TC_CONST_OFF:].view(torch.float32)
    q = f"{cp}.edge_message_fns.prot_pp_prot"
    assert torch.equal(consts[144:272], (sd[q + ".1.to_feats_out.0.bias"].double() * W.TC_PRESCALE).float())
    assert torch.equal(consts[W.TC_C_WH0:W.TC_C_WH0 + 17], sd[q + ".0.Wh"][0].float())
    Whu = (sd[q + ".0.Wh"].double() @ sd[q + ".0.Wu"].double()).float()
    assert torch.equal(consts[W.TC_C_WHU0:W.TC_C_WHU0 + 16], Whu[0])
    # vector image of GVP 1: rows 0..15 = Wh^T, rows 16..31 = (Wh.Wu)^T
    v1 = blob[W.TC_SMALL_OFF + W.TC_VEC_OFF + 2048:W.TC_SMALL_OFF + W.TC_VEC_OFF + 4096]
    rec = decode(v1[:1024], 32) + decode(v1[1024:], 32)
    Wh1, Wu1 = sd[q + ".1.Wh"].double(), sd[q + ".1.Wu"].double()
    want = torch.cat([Wh1.t(), (Wh1 @ Wu1).t()]).float()
    assert torch.allclose(rec, want, rtol=2 ** -21, atol=1e-7)


def _gloo_worker(rank, world, port, sizes, out_dir):
    import os
    import torch
    import torch.distributed as dist
    from pharmacoforge_b200.sharding import gather_results, shard_ranges
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = shard_ranges(sizes, world, pocket_atoms=[400, 250, 1500])[rank]
        flat = [(p, i, n) for p, szs in enumerate(sizes) for i, n in enumerate(szs)]
        # every graph of the shard contributes n rows tagged with its global graph id (stand-in for [x, h] rows)
        rows = [torch.full((flat[gidx][2], 9), float(gidx)) for gidx in rng]
        local = torch.cat(rows) if rows else torch.zeros(0, 9)
        parts = gather_results(local)
        if rank == 0:
            torch.save([p.clone() for p in parts], os.path.join(out_dir, "gathered.pt"))
        else:
            assert parts is None
    finally:
        dist.destroy_process_group()


def test_sharded_gather_world_size_2_gloo(tmp_path):
    """The N>1 sampling path on CPU: contiguous graph shards, no collective except the final gather to rank 0."""
    import socket
    import torch
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    sizes = [[3, 4, 5], [8], [6, 7, 3, 4]]
    mp.spawn(_gloo_worker, args=(2, port, sizes, str(tmp_path)), nprocs=2, join=True)
    parts = torch.load(os.path.join(str(tmp_path), "gathered.pt"))
    got = torch.cat(parts)
    want = torch.cat([torch.full((n, 9), float(g)) for g, n in enumerate(n for szs in sizes for n in szs)])
    assert torch.equal(got, want)      # rank order == graph order: results concatenate without a permutation


def _allreduce_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from pharmacoforge_b200.sharding import allreduce_gradients
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3), torch.nn.Linear(3, 3))  # last one unused
    x = torch.full((4, 5), float(rank + 1))
    net[1](net[0](x)).sum().backward()
    n = allreduce_gradients(net)
    torch.save({"n": n, "g0": net[0].weight.grad.clone(), "dead": net[2].weight.grad}, f"{out_dir}/r{rank}.pt")
    dist.destroy_process_group()


def test_gradient_allreduce_world_size_2_gloo(tmp_path):
    """DDP semantics of the training path on CPU / gloo: one flat all-reduce, mean over ranks, parameters without a
    gradient (dead last-layer protein side) stay grad=None on every rank."""
    import torch.multiprocessing as mp
    port = 29640 + (os.getpid() % 200)
    mp.spawn(_allreduce_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert r0["n"] == r1["n"] == 5 * 7 + 7 + 7 * 3 + 3 + 3 * 3 + 3
    assert torch.equal(r0["g0"], r1["g0"]) and r0["dead"] is None and r1["dead"] is None
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    g = []
    for rank in range(2):
        net.zero_grad()
        net[1](net[0](torch.full((4, 5), float(rank + 1)))).sum().backward()
        g.append(net[0].weight.grad.clone())
    assert torch.allclose(r0["g0"], (g[0] + g[1]) / 2, rtol=1e-6, atol=1e-6)


def test_bench_arguments_name_the_configs(monkeypatch):
    """bench.py: the default is BASELINE.json configs[1]; --workload "configs[2]" selects the 1,500-atom stress shape and
    --precision fp16 the single-pass mode (parsed on the CPU, nothing is launched)."""
    import importlib
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    bench = importlib.import_module("bench")
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.gpus, a.pockets, a.atoms, a.samples, a.precision, a.workload) == (1, 256, 400, 30, "fp32", "configs[1]")
    assert a.warmup >= 3
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "configs[2]", "--precision", "fp16"])
    a = bench.parse()
    assert (a.atoms, a.pockets, a.samples, a.precision) == (1500, 32, 16, "fp16")
    pockets, sizes = None, None
    from pharmacoforge_b200.synthetic import uniform_sizes
    s = uniform_sizes(16, 3, 16, seed=0)
    assert len(s) == 16 and min(s) >= 3 and max(s) <= 16


def test_sort_by_row_groups_edges_in_edge_order():
    """train_ops.sort_by_row: the (perm, ptr) contract of pf_train_gather_bwd_sorted -- row n owns perm[ptr[n]:ptr[n+1]], the
    edges that read row n in ASCENDING edge index (what makes the gather's backward order-fixed), rows without edges are empty."""
    from pharmacoforge_b200.train_ops import sort_by_row
    idx = torch.tensor([4, 1, 4, 0, 1, 4, 6], dtype=torch.int32)
    perm, ptr = sort_by_row(idx, 8)
    assert perm.dtype == torch.int32 and ptr.dtype == torch.int32
    assert ptr.tolist() == [0, 1, 3, 3, 3, 6, 6, 7, 7]
    assert perm.tolist() == [3, 1, 4, 0, 2, 5, 6]
    perm, ptr = sort_by_row(torch.zeros(0, dtype=torch.int32), 3)
    assert perm.numel() == 0 and ptr.tolist() == [0, 0, 0, 0]
