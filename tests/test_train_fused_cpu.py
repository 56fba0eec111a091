"""Host logic of the chain-level training nodes (pharmacoforge_b200/train_fused.py) on CPU: the hand-written tapes of
MessageChain / NodeUpdate / GvpStack against plain autograd of the same composition (gvp.py:89-116, 459-551), with the
CUDA ops replaced by PyTorch restatements of what each kernel computes.  The kernels themselves are checked on the GPU
(tests/test_gpu_training.py); this pins the bookkeeping -- argument order, slices, mask reuse, gradient routing."""
import types

import pytest
import torch


def _gvp_math(feats, vec, Wh, Wu, Wf, bf, Wg, bg, act):
    Vh = torch.einsum("mcv,vh->mch", vec, Wh)
    Vu = torch.einsum("mch,hu->mcu", Vh, Wu)
    sh = torch.sqrt(torch.clamp(Vh.square().sum(1), min=1e-8))
    s = torch.cat([feats, sh], 1)
    z = s @ Wf.t() + bf
    f = torch.nn.functional.silu(z)
    gates = f @ Wg.t() + bg
    vout = (torch.sigmoid(gates) if act else gates)[:, None, :] * Vu
    return f, vout, Vh, Vu, s, z, gates


def _fake_ops():
    T = types.SimpleNamespace()

    def _gvp_fwd(feats, vec, Wh, Wu, Wf, bf, Wg, bg, act):
        f, vout, Vh, Vu, s, z, gates = _gvp_math(feats, vec, Wh, Wu, Wf, bf, Wg, bg, act)
        M = feats.shape[0]
        return f, vout, Vh.reshape(3 * M, -1), Vu.reshape(3 * M, -1), s, z, gates

    def _gvp_bwd(vec, Wh, Wu, Wf, Wg, Vh, Vu, s, z, f, gates, df, dvout, act):
        n = Wf.shape[1] - Wh.shape[1]
        bf, bg = (z - s @ Wf.t())[0], (gates - f @ Wg.t())[0]
        leaves = [t.detach().clone().requires_grad_(True) for t in (s[:, :n], vec, Wh, Wu, Wf, bf, Wg, bg)]
        with torch.enable_grad():
            fo, vo = _gvp_math(*leaves, act)[:2]
            return torch.autograd.grad([fo, vo], leaves, [df, dvout])

    def _layernorm_fwd(x, w, b):
        mu, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
        return (x - mu) / torch.sqrt(var + 1e-5) * w + b, torch.cat([mu, var], 1)

    def layernorm_bwd(x, w, stats, dy):
        leaves = [t.detach().clone().requires_grad_(True) for t in (x, w, torch.zeros_like(w))]
        with torch.enable_grad():
            return torch.autograd.grad(torch.nn.functional.layer_norm(leaves[0], (x.shape[1],), leaves[1], leaves[2], 1e-5),
                                       leaves, dy)

    def _vecln(v):   # gvp.py:163-165 on [rows, 3, U]
        n2 = torch.clamp(v.square().sum(1, keepdim=True), min=1e-8)
        return v / (torch.sqrt(n2.mean(2, keepdim=True) + 1e-5) + 1e-5)

    def vecln_bwd(v, dout):
        leaf = v.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            return torch.autograd.grad(_vecln(leaf), leaf, dout)[0]

    def _segmean(msg, ptr, seg_dst, n_nodes):
        out = torch.zeros((n_nodes,) + tuple(msg.shape[1:]))
        for s_ in range(ptr.numel() - 1):
            a, b = int(ptr[s_]), int(ptr[s_ + 1])
            if b > a:
                d = int(seg_dst[s_]) if seg_dst is not None else s_
                out[d] = out[d] + msg[a:b].mean(0)
        return out

    def segmean_bwd(dout, ptr, seg_dst, n_rows):
        leaf = torch.zeros((n_rows,) + tuple(dout.shape[1:]), requires_grad=True)
        with torch.enable_grad():
            return torch.autograd.grad(_segmean(leaf, ptr, seg_dst, dout.shape[0]), leaf, dout)[0]

    T._gvp_fwd, T._gvp_bwd, T._layernorm_fwd, T.layernorm_bwd = _gvp_fwd, _gvp_bwd, _layernorm_fwd, layernorm_bwd
    T.vecln, T.vecln_bwd, T.segmean, T.segmean_bwd = _vecln, vecln_bwd, _segmean, segmean_bwd
    T.gather = lambda x, idx: x[idx.long()]

    def sort_by_row(idx, n_rows):   # the kernel's contract: edges grouped by the row they read, ascending edge index
        srt = torch.sort(idx.long(), stable=True)
        ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(torch.bincount(srt.values, minlength=n_rows), 0)])
        return srt.indices.to(torch.int32), ptr.to(torch.int32)

    def gather_bwd(dout, idx, n_rows, perm=None, ptr=None):
        if perm is None:
            perm, ptr = sort_by_row(idx, n_rows)
        out = torch.zeros((n_rows,) + tuple(dout.shape[1:]))
        for n in range(n_rows):      # what pf_train_gather_bwd_sorted does: row n adds its edges in perm order
            for e in perm[int(ptr[n]):int(ptr[n + 1])]:
                assert int(idx[int(e)]) == n
                out[n] = out[n] + dout[int(e)]
        return out
    T.gather_bwd, T.sort_by_row = gather_bwd, sort_by_row
    return T


@pytest.fixture()
def F(monkeypatch):
    from pharmacoforge_b200 import train_fused
    monkeypatch.setattr(train_fused, "T", _fake_ops())
    return train_fused


def _gvp_mod(vi, vo, n, no, act=True):
    from pharmacoforge_b200.dynamics import GVP
    return GVP(vi, vo, n, no, vectors_activation=None if act else torch.nn.Identity())


def _ref_gvp(m, feats, vec):
    lin, gl = m.to_feats_out[0], m.scalar_to_vector_gates
    return _gvp_math(feats, vec, m.Wh, m.Wu, lin.weight, lin.bias, gl.weight, gl.bias,
                     isinstance(m.vectors_activation, torch.nn.Sigmoid))[:2]


def _grads(outs, leaves, seeds):
    return torch.autograd.grad([o for o in outs], leaves, seeds, allow_unused=True)


def _close(a, b, what):
    for i, (x, y) in enumerate(zip(a, b)):
        assert (x is None) == (y is None), (what, i)
        if x is not None:
            assert torch.allclose(x, y, rtol=1e-4, atol=1e-5), (what, i, float((x - y).abs().max()))


@pytest.mark.parametrize("with_dst", [False, True])
def test_message_chain_tape_matches_autograd(F, with_dst):
    torch.manual_seed(0)
    ns, nd, E = 9, 7, 20
    gvps = torch.nn.ModuleList([_gvp_mod(17, 16, 144, 128), _gvp_mod(16, 16, 128, 128), _gvp_mod(16, 16, 128, 128)])
    h = torch.randn(ns, 128, requires_grad=True)
    v = torch.randn(ns, 3, 16, requires_grad=True)
    xd, rbf = torch.randn(E, 3), torch.rand(E, 16)
    src = torch.randint(0, ns, (E,), dtype=torch.int32)
    ptr = torch.tensor([0, 3, 3, 8, 12, 20], dtype=torch.int32) if with_dst else torch.tensor([0, 3, 3, 8, 12, 15, 18, 20],
                                                                                               dtype=torch.int32)
    seg_dst = torch.tensor([5, 1, 0, 6, 2], dtype=torch.int32) if with_dst else None
    e = dict(src=src, ptr=ptr, seg_dst=seg_dst, n_dst=nd)
    a_h, a_v = F.message_chain(gvps, h, v, xd, rbf, e)
    T = F.T
    sca, vec = torch.cat([h[src.long()], rbf], 1), torch.cat([xd.unsqueeze(2), v[src.long()]], 2)
    for m in gvps:
        sca, vec = _ref_gvp(m, sca, vec)
    r_h, r_v = T.segmean(sca, ptr, seg_dst, nd), T.segmean(vec, ptr, seg_dst, nd)
    _close([a_h, a_v], [r_h, r_v], "forward")
    leaves = [h, v] + list(gvps.parameters())
    seeds = [torch.randn(nd, 128), torch.randn(nd, 3, 16)]
    _close(_grads([a_h, a_v], leaves, seeds), _grads([r_h, r_v], leaves, seeds), "backward")


@pytest.mark.parametrize("p", [0.0, 0.3])
def test_node_update_tape_matches_autograd(F, p):
    from pharmacoforge_b200.dynamics import GVPMultiEdgeConv, ALL_EDGES
    torch.manual_seed(1)
    conv = GVPMultiEdgeConv(ALL_EDGES, 128, 16, 3, 2, dropout=p)
    N = 11
    ins = [torch.randn(N, 128, requires_grad=True), torch.randn(N, 3, 16, requires_grad=True),
           torch.randn(N, 128, requires_grad=True), torch.randn(N, 3, 16, requires_grad=True)]
    torch.manual_seed(7)
    y, z = F.node_update(conv, "prot", *ins, training=True)
    # reference composition with the same mask draws (fmask, vmask, fmask, vmask)
    torch.manual_seed(7)
    keep = 1.0 - p
    mk = [torch.bernoulli(torch.full(s, keep)) / keep if p > 0 else 1.0
          for s in ((N, 128), (N, 1, 16), (N, 128), (N, 1, 16))]
    T = F.T
    ln1, ln2 = conv.message_layer_norms["prot"].feat_norm, conv.update_layer_norms["prot"].feat_norm
    hh = torch.nn.functional.layer_norm(ins[0] + ins[2] * mk[0], (128,), ln1.weight, ln1.bias, 1e-5)
    vv = T.vecln(ins[1] + ins[3] * mk[1])
    r_h, r_v = hh, vv
    for m in conv.node_update_fns["prot"]:
        r_h, r_v = _ref_gvp(m, r_h, r_v)
    ry = torch.nn.functional.layer_norm(hh + r_h * mk[2], (128,), ln2.weight, ln2.bias, 1e-5)
    rz = T.vecln(vv + r_v * mk[3])
    _close([y, z], [ry, rz], "forward")
    leaves = ins + [ln1.weight, ln1.bias, ln2.weight, ln2.bias] + list(conv.node_update_fns["prot"].parameters())
    seeds = [torch.randn(N, 128), torch.randn(N, 3, 16)]
    _close(_grads([y, z], leaves, seeds), _grads([ry, rz], leaves, seeds), "backward")


def test_gvp_stack_tape_matches_autograd(F):
    torch.manual_seed(2)
    gvps = torch.nn.ModuleList([_gvp_mod(16, 16, 128, 128), _gvp_mod(16, 16, 128, 128), _gvp_mod(16, 1, 128, 64, act=False)])
    s = torch.randn(13, 128, requires_grad=True)
    v = torch.randn(13, 3, 16, requires_grad=True)
    f, vo = F.gvp_stack(gvps, s, v)
    rs, rv = s, v
    for m in gvps:
        rs, rv = _ref_gvp(m, rs, rv)
    _close([f, vo], [rs, rv], "forward")
    leaves = [s, v] + list(gvps.parameters())
    seeds = [torch.randn(13, 64), torch.randn(13, 3, 1)]
    _close(_grads([f, vo], leaves, seeds), _grads([rs, rv], leaves, seeds), "backward")
