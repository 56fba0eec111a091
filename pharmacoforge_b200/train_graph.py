"""Training-mode forward of `PharmRecDynamicsGVP` composed of the differentiable custom ops in `train_ops.py`.

The graph mirrors the reference modules line by line -- GVP.forward (gvp.py:89-116), GVPLayerNorm (gvp.py:159-166),
GVPDropout (gvp.py:121-156), GVPMultiEdgeConv.forward / message (gvp.py:459-551), NoisePredictionBlock
(dynamics_gvp.py:37-42), the encoders and PharmRecDynamicsGVP.forward (dynamics_gvp.py:131-185) -- and reads the
parameters straight from the `nn.Module`s of `dynamics.py`, so `loss.backward()` leaves `.grad` on the reference-named
parameters.  Every arithmetic node is a hand-written CUDA kernel with a hand-written backward; torch does views,
concatenation, residual adds and mask generation.  The per-step graph comes from the same K2 kernel as sampling.
Vectors are component-major [rows, 3, channels].
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

import os

from . import ops
from . import train_fused as F
from . import train_ops as T
from .batch import GraphBatch

# One autograd node per message chain / node update (train_fused.py) instead of one per kernel: same kernels, same
# numbers, ~8x fewer autograd records and op dispatches.  PF_TRAIN_HOST=ops selects the op-by-op graph (A/B, tests).
HOST_FUSED = os.environ.get("PF_TRAIN_HOST", "fused") != "ops"


def gvp_forward(m, feats: torch.Tensor, vec: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """GVP.forward (gvp.py:89-116) on (feats [M, n], vec [M, 3, v]): one differentiable op whose forward / backward
    enqueue the primitive kernels back to back (Vh, Vu, norms, Linear + SiLU, gates, gating)."""
    lin, gl = m.to_feats_out[0], m.scalar_to_vector_gates
    return T.gvp(feats, vec, m.Wh, m.Wu, lin.weight, lin.bias, gl.weight, gl.bias,
                 isinstance(m.vectors_activation, nn.Sigmoid))


def gvp_forward_unfused(m, feats: torch.Tensor, vec: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The same GVP composed of the single-kernel ops (kept for the op-level tests and as the readable definition)."""
    M = feats.shape[0]
    vi, h = m.Wh.shape
    vo = m.Wu.shape[1]
    Vh = T.linear(vec.reshape(M * 3, vi), m.Wh.t().contiguous(), None)       # einsum 'b v c, v h -> b h c'
    Vu = T.linear(Vh, m.Wu.t().contiguous(), None)                           # einsum 'b h c, h u -> b u c'
    sh = T.vecnorm(Vh.view(M, 3, h))
    lin = m.to_feats_out[0]
    f = T.silu(T.linear(torch.cat([feats, sh], dim=1), lin.weight, lin.bias))
    gates = T.linear(f, m.scalar_to_vector_gates.weight, m.scalar_to_vector_gates.bias)
    vout = T.gate(gates, Vu.view(M, 3, vo), isinstance(m.vectors_activation, nn.Sigmoid))
    return f, vout


def gvp_layernorm(m, feats, vec):
    """GVPLayerNorm.forward (gvp.py:159-166)."""
    return T.layernorm(feats, m.feat_norm.weight, m.feat_norm.bias), T.vecln(vec)


def gvp_dropout(m, feats, vec, training: bool):
    """GVPDropout (gvp.py:148-156): nn.Dropout on scalars; whole 3-vectors dropped together (gvp.py:121-146)."""
    p = float(m.feat_dropout.p)
    if not training or p == 0.0:
        return feats, vec
    keep = 1.0 - p
    fmask = torch.bernoulli(torch.full_like(feats, keep)) / keep
    vmask = torch.bernoulli(torch.full((vec.shape[0], 1, vec.shape[2]), keep, device=vec.device)) / keep
    return feats * fmask, vec * vmask


def encoder(seq, x):
    """nn.Sequential(Linear, SiLU, LayerNorm) (dynamics_gvp.py:107-117)."""
    return T.layernorm(T.silu(T.linear(x, seq[0].weight, seq[0].bias)), seq[2].weight, seq[2].bias)


def degree_norms(g: GraphBatch, edges: Dict[str, dict]) -> Dict[str, torch.Tensor]:
    """message_norm = 0 (gvp.py:504-507): per node, (edges of every type into the node type in its graph) / (nodes of the type
    in its graph) + 1.  Per-graph edge counts as add_pharm_edges records them (dynamics_gvp.py:219-221): the pf (= fp) edges
    are assigned through prot_batch_idx[pf_idxs[0]] -- true counts with radius edges, and with kNN edges the edges of
    pharmacophore node i go to the graph that owns protein atom i (the reference's behaviour, see pf_degree_norms)."""
    bi = g.batch_idxs()
    B = g.n_graphs
    count = lambda idx, owner: torch.bincount(owner[idx.long()], minlength=B)
    e_ff = count(edges["ff"]["src"], bi["pharm"])
    e_pp = count(edges["pp"]["dst"], bi["prot"])
    e_pf = count(edges["pf"]["src"] if g.pf_k == 0 else edges["pf"]["dst"], bi["prot"])
    n_f = (g.pharm_ptr[1:] - g.pharm_ptr[:-1]).long()
    n_p = (g.prot_ptr[1:] - g.prot_ptr[:-1]).long()
    return {"pharm": ((e_ff + e_pf) / n_f + 1)[bi["pharm"]], "prot": ((e_pf + e_pp) / n_p + 1)[bi["prot"]]}


def build_edges(g: GraphBatch) -> Dict[str, dict]:
    """Destination-sorted edge lists of the four edge types as int32 tensors: src, dst (per edge), ptr (segment
    offsets into the edge list), seg_dst (node of each segment, None = segment index), n_dst."""
    dev = g.device
    i32 = lambda t: t.to(torch.int32).contiguous()
    k, n = g.pf_k, g.n_pharm
    dyn = g.dynamic_edges()
    out = {}
    pp_dst = torch.repeat_interleave(torch.arange(g.n_prot, device=dev), g.pp_cnt.long())
    out["pp"] = dict(src=i32(g.pp_col), dst=i32(pp_dst), ptr=i32(g.pp_rowptr), seg_dst=None, n_dst=g.n_prot,
                     src_nt="prot", dst_nt="prot", deg=g.pp_cnt[:g.n_prot].float())
    for name, cnt in (("ff", g.ff_cnt), ("pf", g.pf_cnt)):
        ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(cnt[:n].long(), 0)])
        s, d = dyn[name]
        out[name] = dict(src=i32(s), dst=i32(d), ptr=i32(ptr), seg_dst=None, n_dst=n,
                         src_nt="pharm" if name == "ff" else "prot", dst_nt="pharm", deg=cnt[:n].float())
    if k == 0:   # radius pf / fp edges (pf_k == 0): one fp segment per protein atom, identity destinations
        cnt = g.fp_seg_cnt[:g.n_prot]
        ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(cnt.long(), 0)])
        s, d = dyn["fp"]
        out["fp"] = dict(src=i32(s), dst=i32(d), ptr=i32(ptr), seg_dst=None, n_dst=g.n_prot, src_nt="pharm", dst_nt="prot",
                         deg=cnt.float())
        return out
    ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(g.fp_seg_cnt[:k * n].long(), 0)])
    s, d = dyn["fp"]
    fp_deg = torch.zeros(g.n_prot, device=dev).index_add_(0, g.fp_seg_dst[:k * n].long(), g.fp_seg_cnt[:k * n].float())
    out["fp"] = dict(src=i32(s), dst=i32(d), ptr=i32(ptr), seg_dst=i32(g.fp_seg_dst[:k * n]), n_dst=g.n_prot,
                     src_nt="pharm", dst_nt="prot", deg=fp_deg)
    return out


def _norm_scale(conv, e, a_h, a_v):
    """message_norm: 'mean' keeps the per-edge-type mean; a number turns it into SUM / norm = mean * in-degree / norm
    (gvp.py:386-389, 512-517)."""
    nv = getattr(conv, "message_norm", "mean")
    if nv == "mean":
        return a_h, a_v
    sc = e["deg"] / (e["norm0"] if nv == 0 else float(nv))
    return a_h * sc[:, None], a_v * sc[:, None, None]


def conv_forward_fused(conv, feats, edges, geom, training: bool, pharm_only: bool = False):
    """`conv_forward` with one autograd node per edge type and per node type (train_fused.py).  pharm_only: the opt-in
    exact dead-work elimination (dynamics.skip_dead_work) for the LAST layer -- nothing reads its protein side
    (dynamics_gvp.py:84-92), autograd never visits it either, so its pp / fp messages and protein update are not run."""
    agg = {}
    for name in ("ff", "pf", "fp", "pp"):                         # reference etype order (dynamics_gvp.py:46-54)
        e = edges[name]
        if pharm_only and e["dst_nt"] == "prot":
            continue
        key = f"{e['src_nt']}_{name}_{e['dst_nt']}"
        h_src, v_src = feats[e["src_nt"]]
        xd, rbf = geom[name]
        a_h, a_v = _norm_scale(conv, e, *F.message_chain(conv.edge_message_fns[key], h_src, v_src, xd, rbf, e))
        prev = agg.get(e["dst_nt"])
        agg[e["dst_nt"]] = (a_h, a_v) if prev is None else (prev[0] + a_h, prev[1] + a_v)   # cross_reducer="sum"
    return {nt: F.node_update(conv, nt, feats[nt][0], feats[nt][1], agg[nt][0], agg[nt][1], training)
            for nt in (("pharm",) if pharm_only else ("pharm", "prot"))}


def conv_forward(conv, feats, edges, geom, training: bool):
    """GVPMultiEdgeConv.forward (gvp.py:459-538) for message_norm='mean'.  feats[ntype] = (h [N,128], v [N,3,16])."""
    if HOST_FUSED:
        return conv_forward_fused(conv, feats, edges, geom, training)
    agg = {}
    for name in ("ff", "pf", "fp", "pp"):                         # reference etype order (dynamics_gvp.py:46-54)
        e = edges[name]
        key = f"{e['src_nt']}_{name}_{e['dst_nt']}"
        h_src, v_src = feats[e["src_nt"]]
        xd, rbf = geom[name]
        sca = torch.cat([T.gather(h_src, e["src"]), rbf], dim=1)                               # gvp.py:545
        vec = torch.cat([xd.unsqueeze(2), T.gather(v_src, e["src"])], dim=2)                    # gvp.py:543
        for m in conv.edge_message_fns[key]:
            sca, vec = gvp_forward(m, sca, vec)
        a_h = T.segmean(sca, e["ptr"], e["seg_dst"], e["n_dst"])                                # fn.mean, gvp.py:488-497
        a_v = T.segmean(vec, e["ptr"], e["seg_dst"], e["n_dst"])
        a_h, a_v = _norm_scale(conv, e, a_h, a_v)
        if e["dst_nt"] in agg:                                                                  # cross_reducer="sum"
            agg[e["dst_nt"]] = (agg[e["dst_nt"]][0] + a_h, agg[e["dst_nt"]][1] + a_v)
        else:
            agg[e["dst_nt"]] = (a_h, a_v)
    out = {}
    for nt in ("pharm", "prot"):
        h, v = feats[nt]
        m_h, m_v = gvp_dropout(conv.dropout, agg[nt][0], agg[nt][1], training)                  # norm_value = 1.0
        h, v = gvp_layernorm(conv.message_layer_norms[nt], h + m_h, v + m_v)
        r_h, r_v = h, v
        for m in conv.node_update_fns[nt]:
            r_h, r_v = gvp_forward(m, r_h, r_v)
        r_h, r_v = gvp_dropout(conv.dropout, r_h, r_v, training)
        out[nt] = gvp_layernorm(conv.update_layer_norms[nt], h + r_h, v + r_v)
    return out


def dynamics_forward(dyn, g: GraphBatch, t: torch.Tensor, training: bool = True):
    """PharmRecDynamicsGVP.forward (dynamics_gvp.py:131-185) with an autograd graph: (eps_h [Nf, F], eps_x [Nf, 3]).
    g.pharm_x / g.pharm_h hold (x_t, h_t); t [B] is the timestep value per graph."""
    dev = g.device
    if g.pf_k != dyn.pf_k:
        raise ValueError(f"the batch was built for pf_k = {g.pf_k}, the model has pf_k = {dyn.pf_k}")
    if g.pf_k == 0:
        ops.dyn_graph_radius(g.prot_x, g.prot_ptr, g.pharm_x, g.pharm_ptr, float(dyn.graph_cutoffs["ff"]), g.ff_max_nbrs,
                             int(dyn.ff_k), float(dyn.graph_cutoffs["pf"]), g.pf_max_nbrs, g.tile_rows, g.ff_start, g.ff_cnt,
                             g.ff_col, g.pf_start, g.pf_sub_ptr, g.fp_base, g.pf_cnt, g.pf_col, g.pf_sub_start, g.pf_sub_cnt, g.pf_sub_x,
                             g.fp_seg_start, g.fp_seg_cnt, g.fp_col, g.status)
    else:
        ops.dyn_graph(g.prot_x, g.prot_ptr, g.pharm_x, g.pharm_ptr, float(dyn.graph_cutoffs["ff"]), g.ff_max_nbrs, g.pf_k,
                      g.ff_start, g.ff_cnt, g.ff_col, g.pf_cnt, g.pf_col, g.fp_seg_dst, g.fp_seg_start, g.fp_seg_cnt,
                      g.fp_col, g.status, int(dyn.ff_k))
    edges = build_edges(g)
    for name, e in edges.items():   # edges grouped by source row: the deterministic scatter of the gather's backward
        n_src = g.n_prot if e["src_nt"] == "prot" else g.n_pharm
        if name == "pp":            # static graph: sorted once per batch
            if getattr(g, "_pp_src_sorted", None) is None:
                g._pp_src_sorted = T.sort_by_row(e["src"], n_src)
            e["src_sorted"] = g._pp_src_sorted
        else:
            e["src_sorted"] = T.sort_by_row(e["src"], n_src)
    if dyn.message_norm == 0:
        norm0 = degree_norms(g, edges)
        for e in edges.values():
            e["norm0"] = norm0[e["dst_nt"]]
    x = {"pharm": g.pharm_x, "prot": g.prot_x}
    geom = {n: T.edge_geom(x[e["src_nt"]], x[e["dst_nt"]], e["src"], e["dst"]) for n, e in edges.items()}
    bi = g.batch_idxs()
    t = t.to(dev).float()
    h_f = encoder(dyn.pharm_encoder, torch.cat([g.pharm_h, t[bi["pharm"]][:, None]], dim=1).contiguous())
    h_p = encoder(dyn.prot_encoder, torch.cat([g.prot_feats, t[bi["prot"]][:, None]], dim=1).contiguous())
    vs = dyn.vector_size
    feats = {"pharm": (h_f, torch.zeros(g.n_pharm, 3, vs, device=dev)),
             "prot": (h_p, torch.zeros(g.n_prot, 3, vs, device=dev))}
    layers = list(dyn.noise_predictor.conv_layers)
    for li, conv in enumerate(layers):
        if HOST_FUSED and getattr(dyn, "skip_dead_work", False) and li == len(layers) - 1:
            feats = conv_forward_fused(conv, feats, edges, geom, training, pharm_only=True)
        else:
            feats = conv_forward(conv, feats, edges, geom, training)
    head = dyn.noise_predictor.noise_predictor
    sca, vec = feats["pharm"]
    if HOST_FUSED:
        sca, vec = F.gvp_stack(head.gvps, sca, vec)
    else:
        for m in head.gvps:
            sca, vec = gvp_forward(m, sca, vec)
    eps_h = T.linear(sca, head.to_scalar_output.weight, head.to_scalar_output.bias)
    return eps_h, vec.reshape(vec.shape[0], 3)
