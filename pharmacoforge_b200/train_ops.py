"""Differentiable torch.library custom ops of the training path, over the `pf_train_*` entry points of the C ABI.

Each op is one hand-written CUDA kernel (forward) with `register_autograd` wiring its hand-written backward kernel(s);
`train_graph.py` composes the reference's GVP / GVPLayerNorm / GVPMultiEdgeConv graph out of them, so that
`PharmacophoreDiff.forward` in training mode carries an autograd graph whose every node runs in this library.  fp32, CUDA
only, contiguous tensors; vectors are component-major [rows, 3, channels] (the reference is [rows, channels, 3]).
torch itself is used for plumbing only: views, concatenation, residual adds and the final scalar reductions of the loss.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import _lib
from .ops import NS, _f, _i, _s

_L = _lib.load()


_WS = {}


def _workspace(device):
    """The training workspace (see pf_train_workspace_bytes in the header), one per device: partials of every reduction that
    is split over CTAs (split-K weight gradients, bias column sums, LayerNorm affine gradients) and, in its zeroed tail, the
    ticket counters of their deterministic last-CTA reductions.  PF_TRAIN_GEMM=ffma disables the tensor-core kernel AND the
    workspace (A/B switch: every contraction on the fp32 FFMA kernels, split reductions through atomicAdd)."""
    import os
    if os.environ.get("PF_TRAIN_GEMM", "tc") == "ffma":
        return None
    ws = _WS.get(device)
    if ws is None:
        ws = _WS[device] = torch.zeros(_L.pf_train_workspace_bytes() // 4, dtype=torch.float32, device=device)
    return ws


def _raw(t: torch.Tensor) -> int:
    """Device address of a buffer the calling op allocated itself (no validation: see ops._p for the checked form)."""
    return t.data_ptr()


def _ws_args(device):
    ws = _workspace(device)
    return (C.c_void_p(ws.data_ptr()), ws.numel() * 4) if ws is not None else (None, 0)


def _gemm(A, B, bias, Cm, M, N, K, a_rs, a_cs, b_rs, b_cs, accumulate=False, split_k=1):
    _lib.check(_L.pf_train_gemm(_f(A), _f(B), _f(bias), _f(Cm), M, N, K, a_rs, a_cs, b_rs, b_cs, N, int(accumulate),
                                split_k, *_ws_args(A.device), _s()), "pf_train_gemm")


# ------------------------------------------------------------------------------------------------ linear
@torch.library.custom_op(f"{NS}::train_linear", mutates_args=())
def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """y[M,N] = x[M,K] w[N,K]^T + b  (nn.Linear; also the Wh / Wu contractions with b = None)."""
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    _gemm(x, w, b, y, M, N, K, K, 1, 1, K)
    return y


@linear.register_fake
def _(x, w, b):
    return x.new_empty(x.shape[0], w.shape[0])


@torch.library.custom_op(f"{NS}::train_linear_bwd", mutates_args=())
def linear_bwd(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, need_bias: bool) -> Tuple[torch.Tensor, torch.Tensor,
                                                                                            torch.Tensor]:
    M, K = x.shape
    N = w.shape[0]
    dx = torch.empty(M, K, dtype=torch.float32, device=x.device)
    _gemm(dy, w, None, dx, M, K, N, N, 1, K, 1)                       # dx = dy w
    dw = torch.empty(N, K, dtype=torch.float32, device=x.device)
    _gemm(dy, x, None, dw, N, K, M, 1, N, K, 1, split_k=max(1, min(64, M // 512)))   # dw = dy^T x, K' = M split over CTAs
    db = torch.zeros(N if need_bias else 0, dtype=torch.float32, device=x.device)
    if need_bias:
        _lib.check(_L.pf_train_colsum(_f(dy), _f(db), M, N, *_ws_args(dy.device), _s()), "pf_train_colsum")
    return dx, dw, db


@linear_bwd.register_fake
def _(x, w, dy, need_bias):
    return x.new_empty(x.shape), w.new_empty(w.shape), x.new_empty(w.shape[0] if need_bias else 0)


def _linear_setup(ctx, inputs, output):
    x, w, b = inputs
    ctx.save_for_backward(x, w)
    ctx.has_bias = b is not None


def _linear_backward(ctx, dy):
    x, w = ctx.saved_tensors
    dx, dw, db = linear_bwd(x, w, dy.contiguous(), ctx.has_bias)
    return dx, dw, (db if ctx.has_bias else None)


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


# ------------------------------------------------------------------------------------------------ SiLU
@torch.library.custom_op(f"{NS}::train_silu", mutates_args=())
def silu(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    _lib.check(_L.pf_train_silu(_f(x), None, _f(y), x.numel(), _s()), "pf_train_silu")
    return y


@silu.register_fake
def _(x):
    return torch.empty_like(x)


@torch.library.custom_op(f"{NS}::train_silu_bwd", mutates_args=())
def silu_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    dx = torch.empty_like(x)
    _lib.check(_L.pf_train_silu(_f(x), _f(dy), _f(dx), x.numel(), _s()), "pf_train_silu")
    return dx


@silu_bwd.register_fake
def _(x, dy):
    return torch.empty_like(x)


silu.register_autograd(lambda ctx, dy: silu_bwd(ctx.saved_tensors[0], dy.contiguous()),
                       setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


# ------------------------------------------------------------------------------------------------ vector gating
@torch.library.custom_op(f"{NS}::train_gate", mutates_args=())
def gate(g: torch.Tensor, vu: torch.Tensor, act_sigmoid: bool) -> torch.Tensor:
    """out[m,c,u] = act(g[m,u]) * vu[m,c,u]  (gvp.py:108-114)."""
    out = torch.empty_like(vu)
    _lib.check(_L.pf_train_gate(_f(g), _f(vu), None, _f(out), None, g.shape[0], g.shape[1], int(act_sigmoid), _s()),
               "pf_train_gate")
    return out


@gate.register_fake
def _(g, vu, act_sigmoid):
    return torch.empty_like(vu)


@torch.library.custom_op(f"{NS}::train_gate_bwd", mutates_args=())
def gate_bwd(g: torch.Tensor, vu: torch.Tensor, dout: torch.Tensor, act_sigmoid: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    dg, dvu = torch.empty_like(g), torch.empty_like(vu)
    _lib.check(_L.pf_train_gate(_f(g), _f(vu), _f(dout), _f(dg), _f(dvu), g.shape[0], g.shape[1], int(act_sigmoid), _s()),
               "pf_train_gate")
    return dg, dvu


@gate_bwd.register_fake
def _(g, vu, dout, act_sigmoid):
    return torch.empty_like(g), torch.empty_like(vu)


def _gate_setup(ctx, inputs, output):
    g, vu, act = inputs
    ctx.save_for_backward(g, vu)
    ctx.act = act


def _gate_backward(ctx, dout):
    g, vu = ctx.saved_tensors
    dg, dvu = gate_bwd(g, vu, dout.contiguous(), ctx.act)
    return dg, dvu, None


gate.register_autograd(_gate_backward, setup_context=_gate_setup)


# ------------------------------------------------------------------------------------------------ vector norms
@torch.library.custom_op(f"{NS}::train_vecnorm", mutates_args=())
def vecnorm(vh: torch.Tensor) -> torch.Tensor:
    """sh[m,h] = sqrt(max(sum_c vh[m,c,h]^2, 1e-8))  (_norm_no_nan, gvp.py:12-19)."""
    sh = torch.empty(vh.shape[0], vh.shape[2], dtype=torch.float32, device=vh.device)
    _lib.check(_L.pf_train_vecnorm(_f(vh), None, _f(sh), vh.shape[0], vh.shape[2], _s()), "pf_train_vecnorm")
    return sh


@vecnorm.register_fake
def _(vh):
    return vh.new_empty(vh.shape[0], vh.shape[2])


@torch.library.custom_op(f"{NS}::train_vecnorm_bwd", mutates_args=())
def vecnorm_bwd(vh: torch.Tensor, dsh: torch.Tensor) -> torch.Tensor:
    dvh = torch.empty_like(vh)
    _lib.check(_L.pf_train_vecnorm(_f(vh), _f(dsh), _f(dvh), vh.shape[0], vh.shape[2], _s()), "pf_train_vecnorm")
    return dvh


@vecnorm_bwd.register_fake
def _(vh, dsh):
    return torch.empty_like(vh)


vecnorm.register_autograd(lambda ctx, d: vecnorm_bwd(ctx.saved_tensors[0], d.contiguous()),
                          setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


# ------------------------------------------------------------------------------------------------ layer norms
@torch.library.custom_op(f"{NS}::train_layernorm", mutates_args=())
def _layernorm_fwd(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    y = torch.empty_like(x)
    stats = torch.empty(x.shape[0], 2, dtype=torch.float32, device=x.device)
    _lib.check(_L.pf_train_layernorm_fwd(_f(x), _f(w), _f(b), _f(y), _f(stats), x.shape[0], x.shape[1], _s()),
               "pf_train_layernorm_fwd")
    return y, stats


@_layernorm_fwd.register_fake
def _(x, w, b):
    return torch.empty_like(x), x.new_empty(x.shape[0], 2)


@torch.library.custom_op(f"{NS}::train_layernorm_bwd", mutates_args=())
def layernorm_bwd(x: torch.Tensor, w: torch.Tensor, stats: torch.Tensor, dy: torch.Tensor) -> Tuple[torch.Tensor,
                                                                                                   torch.Tensor, torch.Tensor]:
    dx = torch.empty_like(x)
    dw, db = torch.zeros_like(w), torch.zeros_like(w)
    _lib.check(_L.pf_train_layernorm_bwd(_f(x), _f(w), _f(stats), _f(dy), _f(dx), _f(dw), _f(db), x.shape[0], x.shape[1],
                                         *_ws_args(x.device), _s()), "pf_train_layernorm_bwd")
    return dx, dw, db


@layernorm_bwd.register_fake
def _(x, w, stats, dy):
    return torch.empty_like(x), torch.empty_like(w), torch.empty_like(w)


def _ln_setup(ctx, inputs, output):
    x, w, b = inputs
    ctx.save_for_backward(x, w, output[1])


def _ln_backward(ctx, dy, dstats):
    x, w, stats = ctx.saved_tensors
    return layernorm_bwd(x, w, stats, dy.contiguous())


_layernorm_fwd.register_autograd(_ln_backward, setup_context=_ln_setup)


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """nn.LayerNorm over the last dimension, eps 1e-5."""
    return _layernorm_fwd(x, w, b)[0]


@torch.library.custom_op(f"{NS}::train_vecln", mutates_args=())
def vecln(v: torch.Tensor) -> torch.Tensor:
    """Vector half of GVPLayerNorm (gvp.py:163-165) on [rows, 3, U]."""
    out = torch.empty_like(v)
    _lib.check(_L.pf_train_vecln(_f(v), None, _f(out), v.shape[0], v.shape[2], _s()), "pf_train_vecln")
    return out


@vecln.register_fake
def _(v):
    return torch.empty_like(v)


@torch.library.custom_op(f"{NS}::train_vecln_bwd", mutates_args=())
def vecln_bwd(v: torch.Tensor, dout: torch.Tensor) -> torch.Tensor:
    dv = torch.empty_like(v)
    _lib.check(_L.pf_train_vecln(_f(v), _f(dout), _f(dv), v.shape[0], v.shape[2], _s()), "pf_train_vecln")
    return dv


@vecln_bwd.register_fake
def _(v, dout):
    return torch.empty_like(v)


vecln.register_autograd(lambda ctx, d: vecln_bwd(ctx.saved_tensors[0], d.contiguous()),
                        setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


# ------------------------------------------------------------------------------------------------ graph data movement
@torch.library.custom_op(f"{NS}::train_gather", mutates_args=())
def gather(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """out[e] = x[idx[e]] over rows of width D = prod(x.shape[1:])  (edges.src[...], gvp.py:543-545)."""
    D = int(math.prod(x.shape[1:]))
    out = torch.empty((idx.numel(),) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    _lib.check(_L.pf_train_gather(_f(x), _i(idx), _f(out), idx.numel(), D, 0, _s()), "pf_train_gather")
    return out


@gather.register_fake
def _(x, idx):
    return x.new_empty((idx.numel(),) + tuple(x.shape[1:]))


def sort_by_row(idx: torch.Tensor, n_rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(perm, ptr) for the deterministic gather backward: the edges grouped by the row they read (stable sort: ascending
    edge index within a row), row n owns perm[ptr[n] : ptr[n + 1]]."""
    srt = torch.sort(idx.long(), stable=True)
    ptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=idx.device)
    torch.cumsum(torch.bincount(srt.values, minlength=n_rows), 0, out=ptr[1:])
    return srt.indices.to(torch.int32), ptr.to(torch.int32)


@torch.library.custom_op(f"{NS}::train_gather_bwd", mutates_args=())
def gather_bwd(dout: torch.Tensor, idx: torch.Tensor, n_rows: int, perm: Optional[torch.Tensor] = None,
               ptr: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx[n] = sum of dout[e] over the edges e with idx[e] == n, added in ascending edge order (no atomics: gradients are
    bit-reproducible).  (perm, ptr) = sort_by_row(idx, n_rows) when the caller has it (train_graph.build_edges computes it
    once per edge type and step), else it is computed here."""
    if perm is None or ptr is None:
        perm, ptr = sort_by_row(idx, n_rows)
    dx = torch.empty((n_rows,) + tuple(dout.shape[1:]), dtype=torch.float32, device=dout.device)
    D = int(math.prod(dout.shape[1:]))
    _lib.check(_L.pf_train_gather_bwd_sorted(_f(dout), _i(perm), _i(ptr), _f(dx), n_rows, D, _s()),
               "pf_train_gather_bwd_sorted")
    return dx


@gather_bwd.register_fake
def _(dout, idx, n_rows, perm=None, ptr=None):
    return dout.new_empty((n_rows,) + tuple(dout.shape[1:]))


def _gather_setup(ctx, inputs, output):
    x, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n_rows = x.shape[0]


gather.register_autograd(lambda ctx, d: (gather_bwd(d.contiguous(), ctx.saved_tensors[0], ctx.n_rows), None),
                         setup_context=_gather_setup)


@torch.library.custom_op(f"{NS}::train_segmean", mutates_args=())
def segmean(msg: torch.Tensor, ptr: torch.Tensor, seg_dst: Optional[torch.Tensor], n_nodes: int) -> torch.Tensor:
    """Mean of the destination-sorted message rows per segment (fn.mean, gvp.py:488-497): segment s = rows
    [ptr[s], ptr[s+1]) -> node seg_dst[s] (None: s).  Nodes without a segment / with an empty one get zeros."""
    out = torch.zeros((n_nodes,) + tuple(msg.shape[1:]), dtype=torch.float32, device=msg.device)
    D = int(math.prod(msg.shape[1:]))
    _lib.check(_L.pf_train_segmean(_f(msg), _i(ptr), _i(seg_dst), _f(out), ptr.numel() - 1, D, 0, _s()), "pf_train_segmean")
    return out


@segmean.register_fake
def _(msg, ptr, seg_dst, n_nodes):
    return msg.new_empty((n_nodes,) + tuple(msg.shape[1:]))


@torch.library.custom_op(f"{NS}::train_segmean_bwd", mutates_args=())
def segmean_bwd(dout: torch.Tensor, ptr: torch.Tensor, seg_dst: Optional[torch.Tensor], n_rows: int) -> torch.Tensor:
    dmsg = torch.zeros((n_rows,) + tuple(dout.shape[1:]), dtype=torch.float32, device=dout.device)
    D = int(math.prod(dout.shape[1:]))
    _lib.check(_L.pf_train_segmean(_f(dout), _i(ptr), _i(seg_dst), _f(dmsg), ptr.numel() - 1, D, 1, _s()), "pf_train_segmean")
    return dmsg


@segmean_bwd.register_fake
def _(dout, ptr, seg_dst, n_rows):
    return dout.new_empty((n_rows,) + tuple(dout.shape[1:]))


def _segmean_setup(ctx, inputs, output):
    msg, ptr, seg_dst, n_nodes = inputs
    ctx.save_for_backward(ptr, seg_dst) if seg_dst is not None else ctx.save_for_backward(ptr)
    ctx.has_dst = seg_dst is not None
    ctx.n_rows = msg.shape[0]


def _segmean_backward(ctx, dout):
    saved = ctx.saved_tensors
    ptr, seg_dst = saved[0], (saved[1] if ctx.has_dst else None)
    return segmean_bwd(dout.contiguous(), ptr, seg_dst, ctx.n_rows), None, None, None


segmean.register_autograd(_segmean_backward, setup_context=_segmean_setup)


@torch.library.custom_op(f"{NS}::train_edge_geom", mutates_args=())
def edge_geom(src_x: torch.Tensor, dst_x: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> Tuple[torch.Tensor,
                                                                                                      torch.Tensor]:
    """x_diff [E,3] (unit vectors) and rbf [E,16] of every edge (gvp.py:472-480).  Coordinates are data: no gradient."""
    E = src.numel()
    xd = torch.empty(E, 3, dtype=torch.float32, device=src_x.device)
    rbf = torch.empty(E, 16, dtype=torch.float32, device=src_x.device)
    _lib.check(_L.pf_train_edge_geom(_f(src_x), _f(dst_x), _i(src), _i(dst), _f(xd), _f(rbf), E, _s()), "pf_train_edge_geom")
    return xd, rbf


@edge_geom.register_fake
def _(src_x, dst_x, src, dst):
    return src_x.new_empty(src.numel(), 3), src_x.new_empty(src.numel(), 16)


# ------------------------------------------------------------------------------------------------ one GVP per op
@torch.library.custom_op(f"{NS}::train_gvp", mutates_args=())
def _gvp_fwd(feats: torch.Tensor, vec: torch.Tensor, Wh: torch.Tensor, Wu: torch.Tensor, Wf: torch.Tensor,
             bf: torch.Tensor, Wg: torch.Tensor, bg: torch.Tensor, act_sigmoid: bool) -> Tuple[
                 torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    M, n = feats.shape
    vi, h = Wh.shape
    vo, no = Wu.shape[1], Wf.shape[0]
    e = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=feats.device)
    Vh, Vu, s, z, f, gates, vout = e(3 * M, h), e(3 * M, vo), e(M, n + h), e(M, no), e(M, no), e(M, vo), e(M, 3, vo)
    q = _raw   # buffers allocated right here: fp32, contiguous, on the inputs' device -- no need to re-validate them
    _lib.check(_L.pf_train_gvp_fwd(_f(feats), _f(vec), _f(Wh), _f(Wu), _f(Wf), _f(bf), _f(Wg), _f(bg), M, n, vi, h, vo, no,
                                   int(act_sigmoid), q(Vh), q(Vu), q(s), q(z), q(f), q(gates), q(vout),
                                   *_ws_args(feats.device), _s()), "pf_train_gvp_fwd")
    return f, vout, Vh, Vu, s, z, gates


@_gvp_fwd.register_fake
def _(feats, vec, Wh, Wu, Wf, bf, Wg, bg, act_sigmoid):
    M, n = feats.shape
    h, vo, no = Wh.shape[1], Wu.shape[1], Wf.shape[0]
    e = feats.new_empty
    return e(M, no), e(M, 3, vo), e(3 * M, h), e(3 * M, vo), e(M, n + h), e(M, no), e(M, vo)


@torch.library.custom_op(f"{NS}::train_gvp_bwd", mutates_args=())
def _gvp_bwd(vec: torch.Tensor, Wh: torch.Tensor, Wu: torch.Tensor, Wf: torch.Tensor, Wg: torch.Tensor, Vh: torch.Tensor,
             Vu: torch.Tensor, s: torch.Tensor, z: torch.Tensor, f: torch.Tensor, gates: torch.Tensor, df: torch.Tensor,
             dvout: torch.Tensor, act_sigmoid: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                                              torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    M = z.shape[0]
    vi, h = Wh.shape
    vo, no = Wu.shape[1], Wf.shape[0]
    n = Wf.shape[1] - h
    e = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=z.device)
    # the five per-row temporaries (dgates, dVu, dfz, ds, dVh) share one allocation: they never leave this call
    al = lambda k: (k + 63) & ~63
    offs, tot = [], 0
    for k in (M * vo, 3 * M * vo, M * no, M * (n + h), 3 * M * h):
        offs.append(tot)
        tot += al(k)
    tmp = e(max(tot, 1))
    t0 = tmp.data_ptr()
    dgates, dVu, dfz, ds, dVh = (t0 + 4 * o for o in offs)
    dfeats, dvec = e(M, n), e(M, 3, vi)
    dWh, dWu, dWf, dWg = e(vi, h), e(h, vo), e(no, n + h), e(vo, no)
    dbf, dbg = torch.zeros(no, device=z.device), torch.zeros(vo, device=z.device)
    q = _raw
    _lib.check(_L.pf_train_gvp_bwd(_f(vec), _f(Wh), _f(Wu), _f(Wf), _f(Wg), _f(Vh), _f(Vu), _f(s), _f(z), _f(f), _f(gates),
                                   _f(df), _f(dvout), M, n, vi, h, vo, no, int(act_sigmoid), dgates, dVu, dfz,
                                   ds, dVh, q(dfeats), q(dvec), q(dWh), q(dWu), q(dWf), q(dbf), q(dWg),
                                   q(dbg), *_ws_args(z.device), _s()), "pf_train_gvp_bwd")
    return dfeats, dvec, dWh, dWu, dWf, dbf, dWg, dbg


@_gvp_bwd.register_fake
def _(vec, Wh, Wu, Wf, Wg, Vh, Vu, s, z, f, gates, df, dvout, act_sigmoid):
    M = z.shape[0]
    n = Wf.shape[1] - Wh.shape[1]
    e = z.new_empty
    return (e(M, n), e(M, 3, Wh.shape[0]), e(Wh.shape), e(Wu.shape), e(Wf.shape), e(Wf.shape[0]), e(Wg.shape),
            e(Wg.shape[0]))


def _gvp_setup(ctx, inputs, output):
    feats, vec, Wh, Wu, Wf, bf, Wg, bg, act = inputs
    f, vout, Vh, Vu, s, z, gates = output
    ctx.save_for_backward(vec, Wh, Wu, Wf, Wg, Vh, Vu, s, z, f, gates)
    ctx.act = act


def _gvp_backward(ctx, df, dvout, *unused):
    vec, Wh, Wu, Wf, Wg, Vh, Vu, s, z, f, gates = ctx.saved_tensors
    df = torch.zeros_like(f) if df is None else df.contiguous()
    dvout = torch.zeros(z.shape[0], 3, Wu.shape[1], device=z.device) if dvout is None else dvout.contiguous()
    dfeats, dvec, dWh, dWu, dWf, dbf, dWg, dbg = _gvp_bwd(vec, Wh, Wu, Wf, Wg, Vh, Vu, s, z, f, gates, df, dvout, ctx.act)
    return dfeats, dvec, dWh, dWu, dWf, dbf, dWg, dbg, None


_gvp_fwd.register_autograd(_gvp_backward, setup_context=_gvp_setup)


def gvp(feats: torch.Tensor, vec: torch.Tensor, Wh: torch.Tensor, Wu: torch.Tensor, Wf: torch.Tensor, bf: torch.Tensor,
        Wg: torch.Tensor, bg: torch.Tensor, act_sigmoid: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """GVP.forward (gvp.py:89-116) as one differentiable op: (feats [M,n], vec [M,3,vi]) -> (f [M,no], vout [M,3,vo])."""
    out = _gvp_fwd(feats.contiguous(), vec.contiguous(), Wh, Wu, Wf, bf, Wg, bg, act_sigmoid)
    return out[0], out[1]
