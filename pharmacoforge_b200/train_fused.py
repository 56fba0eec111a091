"""Chain-level autograd nodes of the training path.

`train_graph.py` used to hand every GVP, gather, concatenation, LayerNorm and mean to autograd as its own node: ~400
`torch.library` op calls with an autograd record each per step (forward + backward), 35 ms of host time against ~20 ms
of GPU time.  Here one `torch.autograd.Function` covers a whole edge-type message chain
(gather -> 3 GVPs -> segmented means, gvp.py:540-551 + 488-497) or a whole node update (dropout -> residual ->
GVPLayerNorm -> 2 GVPs -> dropout -> residual -> GVPLayerNorm, gvp.py:511-532) with a hand-written tape.  The kernels are
the same ones: every arithmetic step still goes through the registered `pharmacoforge::train_*` custom ops of
`train_ops.py` (forward *and* backward kernels, called without autograd recording from inside the Function), so the
numbers are identical to the op-by-op graph (`tests/test_gpu_training.py::test_fused_host_graph_matches_op_by_op_graph`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import train_ops as T

_GVP_NP = 6  # parameters per GVP: Wh, Wu, Wf, bf, Wg, bg


def gvp_params(m) -> Tuple[torch.Tensor, ...]:
    lin, gl = m.to_feats_out[0], m.scalar_to_vector_gates
    return (m.Wh, m.Wu, lin.weight, lin.bias, gl.weight, gl.bias)


def _gvp_chain_fwd(sca, vec, params: Sequence[torch.Tensor], acts: Sequence[bool]):
    """GVP x n on (sca, vec): returns the outputs and, per GVP, what its backward needs."""
    tape = []
    for i, act in enumerate(acts):
        Wh, Wu, Wf, bf, Wg, bg = params[_GVP_NP * i:_GVP_NP * (i + 1)]
        f, vout, Vh, Vu, s, z, gates = T._gvp_fwd(sca, vec, Wh, Wu, Wf, bf, Wg, bg, act)
        tape.append((vec, Vh, Vu, s, z, f, gates))
        sca, vec = f, vout
    return sca, vec, tape


def _gvp_chain_bwd(df, dv, params: Sequence[torch.Tensor], acts: Sequence[bool], tape):
    """Reverse of `_gvp_chain_fwd`: (d sca_in, d vec_in, parameter gradients in `params` order)."""
    grads: List[Optional[torch.Tensor]] = [None] * len(params)
    for i in range(len(acts) - 1, -1, -1):
        Wh, Wu, Wf, bf, Wg, bg = params[_GVP_NP * i:_GVP_NP * (i + 1)]
        vec, Vh, Vu, s, z, f, gates = tape[i]
        df, dv, dWh, dWu, dWf, dbf, dWg, dbg = T._gvp_bwd(vec, Wh, Wu, Wf, Wg, Vh, Vu, s, z, f, gates, df, dv, acts[i])
        grads[_GVP_NP * i:_GVP_NP * (i + 1)] = [dWh, dWu, dWf, dbf, dWg, dbg]
    return df, dv, grads


def _flatten_tape(tape):
    return [t for rec in tape for t in rec]


def _unflatten_tape(flat, n):
    return [tuple(flat[7 * i:7 * (i + 1)]) for i in range(n)]


class MessageChain(torch.autograd.Function):
    """One edge type of GVPMultiEdgeConv.forward: messages of all edges (gvp.py:540-551 -> 89-116) and their mean per
    destination (fn.mean, gvp.py:488-497).  (h_src [Ns,128], v_src [Ns,3,16]) -> (a_h [Nd,128], a_v [Nd,3,16])."""

    @staticmethod
    def forward(ctx, h_src, v_src, xd, rbf, src, ptr, seg_dst, n_dst: int, acts: Tuple[bool, ...], src_sorted, *params):
        # src_sorted = T.sort_by_row(src, n_src) or None: the edges grouped by source row, for the deterministic scatter of
        # the gather's backward (computed once per edge type and step by train_graph.build_edges)
        gh, gv = T.gather(h_src, src), T.gather(v_src, src)
        sca = torch.cat([gh, rbf], dim=1)                                   # gvp.py:545
        vec = torch.cat([xd.unsqueeze(2), gv], dim=2)                       # gvp.py:543
        sca, vec, tape = _gvp_chain_fwd(sca, vec, params, acts)
        a_h = T.segmean(sca, ptr, seg_dst, n_dst)
        a_v = T.segmean(vec, ptr, seg_dst, n_dst)
        ctx.acts, ctx.n_src, ctx.n_edges, ctx.has_dst = acts, h_src.shape[0], src.numel(), seg_dst is not None
        ctx.n_h = h_src.shape[1]
        ctx.src_sorted = src_sorted
        idx = (src, ptr, seg_dst) if seg_dst is not None else (src, ptr)
        ctx.save_for_backward(*idx, *params, *_flatten_tape(tape))
        ctx.n_idx, ctx.n_params = len(idx), len(params)
        return a_h, a_v

    @staticmethod
    def backward(ctx, da_h, da_v):
        saved = ctx.saved_tensors
        idx, params = saved[:ctx.n_idx], saved[ctx.n_idx:ctx.n_idx + ctx.n_params]
        tape = _unflatten_tape(saved[ctx.n_idx + ctx.n_params:], len(ctx.acts))
        src, ptr = idx[0], idx[1]
        seg_dst = idx[2] if ctx.has_dst else None
        f_last, vout_shape = tape[-1][5], (ctx.n_edges, 3, params[-5].shape[1])   # Wu of the last GVP: [h, vo]
        df = (T.segmean_bwd(da_h.contiguous(), ptr, seg_dst, ctx.n_edges) if da_h is not None
              else torch.zeros_like(f_last))
        dv = (T.segmean_bwd(da_v.contiguous(), ptr, seg_dst, ctx.n_edges) if da_v is not None
              else torch.zeros(vout_shape, device=f_last.device))
        df, dv, grads = _gvp_chain_bwd(df, dv, params, ctx.acts, tape)
        dh_src = dv_src = None
        perm, sptr = ctx.src_sorted if ctx.src_sorted is not None else T.sort_by_row(src, ctx.n_src)
        if ctx.needs_input_grad[0]:
            dh_src = T.gather_bwd(df[:, :ctx.n_h].contiguous(), src, ctx.n_src, perm, sptr)
        if ctx.needs_input_grad[1]:
            dv_src = T.gather_bwd(dv[:, :, 1:].contiguous(), src, ctx.n_src, perm, sptr)
        return (dh_src, dv_src, None, None, None, None, None, None, None, None, *grads)


class NodeUpdate(torch.autograd.Function):
    """The node half of GVPMultiEdgeConv.forward for one node type (gvp.py:511-532):
    (h, v) <- GVPLayerNorm_msg(h + drop(m_h), v + drop(m_v)); (r_h, r_v) = GVP x n; (h, v) <- GVPLayerNorm_upd(h + drop(r_h),
    v + drop(r_v)).  Dropout masks (scaled by 1 / keep; whole 3-vectors dropped together, gvp.py:121-146) are drawn here."""

    @staticmethod
    def forward(ctx, h, v, m_h, m_v, p_drop: float, acts: Tuple[bool, ...], w1, b1, w2, b2, *params):
        masks: List[Optional[torch.Tensor]] = [None, None, None, None]
        if p_drop > 0.0:
            keep = 1.0 - p_drop
            for i in range(4):      # same draws, in the same order, as train_graph.gvp_dropout on (m_h, m_v), (r_h, r_v)
                shape = h.shape if i % 2 == 0 else (v.shape[0], 1, v.shape[2])
                masks[i] = torch.bernoulli(torch.full(shape, keep, device=h.device)) / keep
            m_h, m_v = m_h * masks[0], m_v * masks[1]
        x1, u1 = h + m_h, v + m_v
        y1, st1 = T._layernorm_fwd(x1, w1, b1)
        z1 = T.vecln(u1)
        r_h, r_v, tape = _gvp_chain_fwd(y1, z1, params, acts)
        if p_drop > 0.0:
            r_h, r_v = r_h * masks[2], r_v * masks[3]
        x2, u2 = y1 + r_h, z1 + r_v
        y2, st2 = T._layernorm_fwd(x2, w2, b2)
        z2 = T.vecln(u2)
        ctx.acts, ctx.has_masks = acts, p_drop > 0.0
        mk = tuple(masks) if ctx.has_masks else ()
        ctx.save_for_backward(x1, u1, st1, x2, u2, st2, w1, w2, *mk, *params, *_flatten_tape(tape))
        ctx.n_params = len(params)
        return y2, z2

    @staticmethod
    def backward(ctx, dy2, dz2):
        saved = ctx.saved_tensors
        x1, u1, st1, x2, u2, st2, w1, w2 = saved[:8]
        o = 8
        masks = saved[o:o + 4] if ctx.has_masks else (None,) * 4
        o += 4 if ctx.has_masks else 0
        params = saved[o:o + ctx.n_params]
        tape = _unflatten_tape(saved[o + ctx.n_params:], len(ctx.acts))
        dy2 = torch.zeros_like(x2) if dy2 is None else dy2.contiguous()
        dz2 = torch.zeros_like(u2) if dz2 is None else dz2.contiguous()
        dx2, dw2, db2 = T.layernorm_bwd(x2, w2, st2, dy2)
        du2 = T.vecln_bwd(u2, dz2)
        dr_h, dr_v = (dx2 * masks[2], du2 * masks[3]) if ctx.has_masks else (dx2, du2)
        dfe, dve, grads = _gvp_chain_bwd(dr_h.contiguous(), dr_v.contiguous(), params, ctx.acts, tape)
        dy1, dz1 = dx2 + dfe, du2 + dve
        dx1, dw1, db1 = T.layernorm_bwd(x1, w1, st1, dy1)
        du1 = T.vecln_bwd(u1, dz1)
        dm_h, dm_v = (dx1 * masks[0], du1 * masks[1]) if ctx.has_masks else (dx1, du1)
        return (dx1, du1, dm_h, dm_v, None, None, dw1, db1, dw2, db2, *grads)


class GvpStack(torch.autograd.Function):
    """GVP x n as one node (the noise head's chain, dynamics_gvp.py:37-42)."""

    @staticmethod
    def forward(ctx, sca, vec, acts: Tuple[bool, ...], *params):
        f, vout, tape = _gvp_chain_fwd(sca.contiguous(), vec.contiguous(), params, acts)
        ctx.acts, ctx.n_params = acts, len(params)
        ctx.save_for_backward(*params, *_flatten_tape(tape))
        return f, vout

    @staticmethod
    def backward(ctx, df, dv):
        saved = ctx.saved_tensors
        params = saved[:ctx.n_params]
        tape = _unflatten_tape(saved[ctx.n_params:], len(ctx.acts))
        f_last = tape[-1][5]
        df = torch.zeros_like(f_last) if df is None else df.contiguous()
        dv = (torch.zeros(f_last.shape[0], 3, params[-5].shape[1], device=f_last.device) if dv is None
              else dv.contiguous())
        dfe, dve, grads = _gvp_chain_bwd(df, dv, params, ctx.acts, tape)
        return (dfe, dve, None, *grads)


def _acts(gvps) -> Tuple[bool, ...]:
    return tuple(isinstance(m.vectors_activation, nn.Sigmoid) for m in gvps)


def _params(gvps) -> Tuple[torch.Tensor, ...]:
    return tuple(p for m in gvps for p in gvp_params(m))


def message_chain(gvps, h_src, v_src, xd, rbf, e):
    return MessageChain.apply(h_src, v_src, xd, rbf, e["src"], e["ptr"], e["seg_dst"], e["n_dst"], _acts(gvps),
                              e.get("src_sorted"), *_params(gvps))


def node_update(conv, nt: str, h, v, m_h, m_v, training: bool):
    p = float(conv.dropout.feat_dropout.p) if training else 0.0
    gvps = conv.node_update_fns[nt]
    ln1, ln2 = conv.message_layer_norms[nt].feat_norm, conv.update_layer_norms[nt].feat_norm
    return NodeUpdate.apply(h, v, m_h, m_v, p, _acts(gvps), ln1.weight, ln1.bias, ln2.weight, ln2.bias, *_params(gvps))


def gvp_stack(gvps, sca, vec):
    return GvpStack.apply(sca, vec, _acts(gvps), *_params(gvps))
