"""tcgen05 building blocks on the B200: descriptor encodings, TMEM operand layout, bulk copy + mbarrier protocol."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (32, 128), (128, 16), (144, 128), (176, 128)])
@pytest.mark.parametrize("npass", [1, 3])
def test_tc_gemm_selftest(K, N, npass):
    from pharmacoforge_b200 import _lib
    from pharmacoforge_b200.weights import split_bf16, umma_b_image
    lib = _lib.load()
    gen = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=gen)
    W = torch.randn(N, K, generator=gen) / K ** 0.5
    hi, lo = split_bf16(W)
    img_hi, img_lo = umma_b_image(hi).cuda(), umma_b_image(lo).cuda()
    Ad = A.cuda().contiguous()
    D = torch.full((128, N), float("nan"), device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pf_tc_selftest(C.c_void_p(Ad.data_ptr()), C.c_void_p(img_hi.data_ptr()),
                                  C.c_void_p(img_lo.data_ptr()), C.c_void_p(D.data_ptr()), K, N, npass, s),
               "pf_tc_selftest")
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t())
    err = (D.cpu().double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = 2e-5 if npass == 3 else 2e-2
    assert err <= tol * scale, f"K={K} N={N} npass={npass}: max abs err {err:.3e} (scale {scale:.2f})"
    if npass == 3:   # the compensated product must be far better than a single bf16 pass
        assert err <= 1e-3 * scale


def _tc_gemm(A, B, bias, C, M, N, K, a_rs, a_cs, b_rs, b_cs, accumulate=False):
    from pharmacoforge_b200 import _lib
    lib = _lib.load()
    ws = torch.empty(lib.pf_tc_gemm_workspace_bytes(128, 176, 0) // 4, dtype=torch.float32, device="cuda")
    vp = lambda t: C_.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.pf_tc_gemm(vp(A), vp(B), vp(bias), vp(C), M, N, K, a_rs, a_cs, b_rs, b_cs, N, int(accumulate), vp(ws),
                              ws.numel() * 4, C_.c_void_p(torch.cuda.current_stream().cuda_stream)), "pf_tc_gemm")
    torch.cuda.synchronize()


C_ = C


@pytest.mark.parametrize("M,K,N", [(5000, 161, 128), (4096, 144, 128), (1025, 128, 161), (3000, 17, 17), (2500, 16, 32),
                                   (1300, 128, 16), (128, 176, 176)])
def test_tc_gemm_forward_and_dgrad_shapes(M, K, N):
    """pf_tc_gemm, resident-B mode (training forward / input gradients): y = x W^T + b and dx = dy W on the tensor cores with
    the bf16 hi/lo split (3 passes, fp32 exponent range), against an fp64 product: error ~2^-16 of the row scale, far
    inside the 1e-3 bar of the training parity tests; tiny operands (1e-7, gradient-sized) keep that relative accuracy."""
    gen = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=gen).cuda()
    w = (torch.randn(N, K, generator=gen) / K ** 0.5).cuda()
    b = torch.randn(N, generator=gen).cuda()
    y = torch.full((M, N), float("nan"), device="cuda")
    _tc_gemm(x, w, b, y, M, N, K, K, 1, 1, K)                       # B(k, n) = w[n, k]
    ref = x.double() @ w.double().t() + b.double()
    scale = ref.abs().max().item()
    assert (y.double() - ref).abs().max().item() <= 3e-5 * scale + 1e-6
    tiny = torch.full((M, N), float("nan"), device="cuda")
    _tc_gemm(x * 1e-7, w, None, tiny, M, N, K, K, 1, 1, K)           # gradient-sized operand: no fp16 range problem
    ref_t = (x.double() * 1e-7) @ w.double().t()
    assert (tiny.double() - ref_t).abs().max().item() <= 3e-5 * ref_t.abs().max().item()
    # input gradient: dx[M, K] = dy[M, N] w[N, K]  (B(k', n') = w[k', n'], row-major) accumulated onto a base
    dy = torch.randn(M, N, generator=gen).cuda()
    base = torch.randn(M, K, generator=gen).cuda()
    dx = base.clone()
    _tc_gemm(dy, w, None, dx, M, K, N, N, 1, K, 1, accumulate=True)
    ref = base.double() + dy.double() @ w.double()
    assert (dx.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("rows,N,K", [(128, 161, 50000), (128, 144, 4096), (16, 128, 33333), (17, 16, 70001)])
def test_tc_gemm_weight_gradient_is_accurate_and_deterministic(rows, N, K):
    """pf_tc_gemm, split-K mode (training weight gradients): dW[rows, N] = dy^T x with the contraction over K edges split
    across CTAs and reduced in CTA order: bit-identical run to run (the FFMA path used atomics), error ~1e-5 of the result
    scale (bf16 split + the tensor core's truncating accumulation over <= 128 K-steps per accumulator)."""
    gen = torch.Generator().manual_seed(rows * 7 + N)
    dy = torch.randn(K, rows, generator=gen).cuda()                 # A(m, k) = dy[k, m]: a_rs = 1, a_cs = rows
    x = torch.randn(K, N, generator=gen).cuda()                     # B(k, n) = x[k, n]:  b_rs = N, b_cs = 1
    outs = []
    for _ in range(2):
        dw = torch.full((rows, N), float("nan"), device="cuda")
        _tc_gemm(dy, x, None, dw, rows, N, K, 1, rows, N, 1)
        outs.append(dw)
    assert torch.equal(outs[0], outs[1])
    ref = dy.double().t() @ x.double()
    err = (outs[0].double() - ref).abs().max().item()
    assert err <= 5e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_mapped_kernels_equal_the_materialised_rows():
    """pf_edge_conv_tc_mapped / pf_node_update_tc_mapped (source / input scalars read through a row map from a small
    table) against the plain kernels on the materialised rows table[map]: bit-identical outputs."""
    import json
    import os
    from pharmacoforge_b200 import ops
    from pharmacoforge_b200.batch import GraphBatch, Pocket
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    from pharmacoforge_b200.hostutil import polynomial_gamma
    from pharmacoforge_b200.synthetic import make_pocket, synth_state_dict
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    layout = json.load(open(os.path.join(root, "tests/golden/state_dict_layout.json")))
    sd = synth_state_dict(layout, seed=0)
    sd["gamma.gamma"] = polynomial_gamma(100, 1e-5, 2.0)
    dyn = dict(vector_size=16, n_convs=2, n_hidden_scalars=128, message_norm="mean", dropout=0.1, ff_k=0, pf_k=5,
               n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4)
    model = PharmacophoreDiff(6, 11, list("abcdef"), n_timesteps=100, graph_config={"graph_cutoffs": {"pp": 3.5, "pf": 8, "fp": 8, "ff": 9}},
                              dynamics_config=dyn, precision=1e-5)
    model.load_state_dict(sd)
    dev = torch.device("cuda:0")
    g = GraphBatch.from_pockets([Pocket.from_numpy(*make_pocket(n, seed=s)) for n, s in ((300, 1), (77, 2), (1, 3))],
                                [[3, 8], [5], [4, 4, 4]], dev)
    W = model.dynamics.packed_weights(dev)
    gen = torch.Generator().manual_seed(5)
    table = torch.randn(g.n_graphs * 11, 128, generator=gen).to(dev)
    row = torch.randint(0, table.shape[0], (g.n_prot,), generator=gen).to(dev).to(torch.int32)
    full = table[row.long()].contiguous()
    # K3 over the pp edges, general kernel without source vectors
    blob = W.tc[3 * W.tc_stride:4 * W.tc_stride]
    outs = []
    for mapped in (False, True):
        ah, av = torch.full((g.n_prot, 128), float("nan"), device=dev), torch.full((g.n_prot, 48), float("nan"), device=dev)
        if mapped:
            ops.edge_conv_tc_mapped(table, row, None, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles,
                                    g.pp_n_tiles, blob, ah, av, False, False)
        else:
            ops.edge_conv_tc(full, None, g.prot_x, g.prot_x, g.pp_start, g.pp_cnt, None, g.pp_col, g.pp_tiles, g.pp_n_tiles,
                             blob, ah, av, False, False)
        outs.append((ah, av))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert bool(torch.isfinite(outs[1][0]).all())
    # K4, first layer (no input vectors)
    agg_h, agg_v = outs[0]
    res = []
    for mapped in (False, True):
        oh, ov = torch.empty(g.n_prot, 128, device=dev), torch.empty(g.n_prot, 48, device=dev)
        if mapped:
            ops.node_update_tc_mapped(table, row, agg_h, agg_v, W.tcu_view(0, 1), oh, ov, False)
        else:
            ops.node_update_tc(full, None, agg_h, agg_v, W.tcu_view(0, 1), oh, ov, False)
        res.append((oh, ov))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert bool(torch.isfinite(res[1][0]).all())
