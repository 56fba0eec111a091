"""CPU oracle for PharmacoForge's denoising hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A flat, functional restatement (plain torch CPU ops, fp32) of the algorithm the
reference implements across pharmacoforge/models/{pharmacodiff,dynamics_gvp,gvp}.py,
pharmacoforge/utils/unorganized_utils.py and the un-vendored torch_cluster / DGL
kernels those files call.  It exists so that the CUDA path can be checked on a
box where /root/reference does not exist.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; the product
package never does.

Parity pin: the reference ships no tests and no golden vectors,
and its third-party kernels are absent, so the pin is the reference's OWN model
code executed here, unmodified, over the shim packages in oracle/shims
(oracle/make_golden.py -> tests/golden/*.npz).  tests/test_oracle_golden.py checks
every function below against those fixtures.  The torch_cluster / DGL semantics
restated in the shims themselves (edge membership rules, mean aggregation) have
no upstream artefact to pin against: that boundary is "parity unpinned".

All tensors are torch CPU float32 / int64.  Weights are addressed by the
reference's state_dict keys.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

ETYPES = (("pharm", "ff", "pharm"), ("prot", "pf", "pharm"), ("pharm", "fp", "prot"), ("prot", "pp", "prot"))
RBF_DMAX = 15.0   # GVPMultiEdgeConv default, gvp.py:350
RBF_DIM = 16


# --------------------------------------------------------------------------- noise schedule
def gamma_table(n_timesteps: int, precision: float, power: float = 2.0) -> torch.Tensor:
    """pharmacodiff.py:602-632 (clip_noise_schedule, polynomial_schedule) and :636-664."""
    steps = n_timesteps + 1
    x = np.linspace(0, steps, steps)
    a2 = (1.0 - np.power(x / steps, power)) ** 2
    a2 = np.concatenate([np.ones(1), a2], axis=0)
    ratio = np.clip(a2[1:] / a2[:-1], a_min=0.001, a_max=1.0)
    a2 = np.cumprod(ratio, axis=0)
    a2 = (1.0 - 2.0 * precision) * a2 + precision
    g = -(np.log(a2) - np.log(1.0 - a2))
    return torch.from_numpy(g).float()


def gamma_at(gamma: torch.Tensor, t: torch.Tensor, n_timesteps: int) -> torch.Tensor:
    """pharmacodiff.py:666-668."""
    return gamma[torch.round(t * n_timesteps).long()]


def posterior_coefficients(gamma_s: torch.Tensor, gamma_t: torch.Tensor):
    """pharmacodiff.py:140-160 and :390-400 -> (alpha_t_given_s, var_terms, sigma_q)."""
    sigma2_ts = -torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t))
    log_a2_t = F.logsigmoid(-gamma_t)
    log_a2_s = F.logsigmoid(-gamma_s)
    alpha_ts = torch.exp(0.5 * (log_a2_t - log_a2_s))
    sigma_ts = torch.sqrt(sigma2_ts)
    sigma_s = torch.sqrt(torch.sigmoid(gamma_s))
    sigma_t = torch.sqrt(torch.sigmoid(gamma_t))
    var_terms = sigma2_ts / alpha_ts / sigma_t
    sigma_q = sigma_ts * sigma_s / sigma_t
    return alpha_ts, var_terms, sigma_q


def endpoint_coefficients(gamma_s: torch.Tensor, gamma_t: torch.Tensor):
    """The endpoint-parameterisation branch of sample_p_zs_given_zt (pharmacodiff.py:413-418):
    mu = c1 z_t + c2 pred with c1 = alpha_{t|s} sigma_s^2 / sigma_t^2, c2 = alpha_s sigma^2_{t|s} / sigma_t^2,
    evaluated with the reference's operator order."""
    sigma2_ts = -torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t))
    log_a2_t = F.logsigmoid(-gamma_t)
    log_a2_s = F.logsigmoid(-gamma_s)
    alpha_ts = torch.exp(0.5 * (log_a2_t - log_a2_s))
    alpha_s = torch.exp(0.5 * log_a2_s)
    sigma_s = torch.sqrt(torch.sigmoid(gamma_s))
    sigma_t = torch.sqrt(torch.sigmoid(gamma_t))
    return alpha_ts * (sigma_s ** 2) / (sigma_t ** 2), alpha_s * sigma2_ts / (sigma_t ** 2)


# --------------------------------------------------------------------------- graph construction
def _sqdist(q: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """Canonical fp32 squared distance ((dx*dx + dy*dy) + dz*dz), no FMA."""
    dx = q[:, None, 0] - c[None, :, 0]
    dy = q[:, None, 1] - c[None, :, 1]
    dz = q[:, None, 2] - c[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def radius_edges(x: torch.Tensor, ptr: torch.Tensor, r: float, max_num_neighbors: int):
    """radius_graph within each segment of `ptr` (protein_pharm_dataset.py:235, dynamics_gvp.py:196).

    Returns (src, dst) int64 sorted by (dst, src): src = neighbour, dst = centre, strict `< r*r`,
    no self loops, first `max_num_neighbors` neighbours in ascending index kept per centre.
    """
    srcs, dsts = [], []
    r2 = torch.tensor(float(r), dtype=torch.float32) * torch.tensor(float(r), dtype=torch.float32)
    for g in range(ptr.numel() - 1):
        a, b = int(ptr[g]), int(ptr[g + 1])
        if b - a < 2:
            continue
        d = _sqdist(x[a:b], x[a:b])
        hit = d < r2
        # torch_cluster asks for max+1 hits including the centre itself, then drops the self pair
        rank = torch.cumsum(hit.to(torch.int64), dim=1)
        hit = hit & (rank <= max_num_neighbors + 1)
        hit.fill_diagonal_(False)
        ci, ni = torch.nonzero(hit, as_tuple=True)
        dsts.append(ci + a)
        srcs.append(ni + a)
    if not srcs:
        z = torch.zeros(0, dtype=torch.int64)
        return z, z.clone()
    return torch.cat(srcs), torch.cat(dsts)


def knn_edges(prot_x, prot_ptr, pharm_x, pharm_ptr, k: int):
    """knn(prot, pharm, k) per graph (dynamics_gvp.py:202): returns (pharm_idx, prot_idx) int64,
    k rows per pharm node in ascending distance, ties to the lower prot index."""
    qs, cs = [], []
    for g in range(prot_ptr.numel() - 1):
        a, b = int(prot_ptr[g]), int(prot_ptr[g + 1])
        qa, qb = int(pharm_ptr[g]), int(pharm_ptr[g + 1])
        if b == a or qb == qa:
            continue
        d = _sqdist(pharm_x[qa:qb], prot_x[a:b])
        kk = min(k, b - a)
        order = torch.sort(d, dim=1, stable=True).indices[:, :kk]
        qs.append(torch.arange(qa, qb)[:, None].expand(-1, kk).reshape(-1))
        cs.append(order.reshape(-1) + a)
    if not qs:
        z = torch.zeros(0, dtype=torch.int64)
        return z, z.clone()
    return torch.cat(qs), torch.cat(cs)


def knn_graph_edges(x: torch.Tensor, ptr: torch.Tensor, k: int):
    """knn_graph(x, k, batch) of torch_cluster (dynamics_gvp.py:194): per centre the k + 1 nearest nodes of its graph
    INCLUDING itself, ordered by (distance, index) (stable sort), with the self pair dropped.  Returns (src = neighbour,
    dst = centre) int64."""
    srcs, dsts = [], []
    for g in range(ptr.numel() - 1):
        a, b = int(ptr[g]), int(ptr[g + 1])
        if b - a < 2:
            continue
        d = _sqdist(x[a:b], x[a:b])
        kk = min(k + 1, b - a)
        order = torch.sort(d, dim=1, stable=True).indices[:, :kk]
        ci = torch.arange(b - a)[:, None].expand(-1, kk).reshape(-1)
        ni = order.reshape(-1)
        m = ci != ni
        dsts.append(ci[m] + a)
        srcs.append(ni[m] + a)
    if not srcs:
        z = torch.zeros(0, dtype=torch.int64)
        return z, z.clone()
    return torch.cat(srcs), torch.cat(dsts)


def batch_index(ptr: torch.Tensor) -> torch.Tensor:
    """unorganized_utils.py:83-95: arange(B).repeat_interleave(nodes per graph)."""
    return torch.arange(ptr.numel() - 1).repeat_interleave(ptr[1:] - ptr[:-1])


# --------------------------------------------------------------------------- GVP building blocks
def norm_no_nan(x: torch.Tensor, keepdims=False, sqrt=True):
    """gvp.py:12-19."""
    out = torch.clamp(torch.sum(torch.square(x), -1, keepdims), min=1e-8)
    return torch.sqrt(out) if sqrt else out


def rbf(d: torch.Tensor) -> torch.Tensor:
    """gvp.py:26-41 with D_min=0, D_max=15, D_count=16 (centres 0..15, width 15/16)."""
    mu = torch.linspace(0.0, RBF_DMAX, RBF_DIM).view(1, -1)
    sig = (RBF_DMAX - 0.0) / RBF_DIM
    return torch.exp(-((d.unsqueeze(-1) - mu) / sig) ** 2)


def gvp(sd: Dict[str, torch.Tensor], p: str, s: torch.Tensor, v: torch.Tensor, vec_act: str = "sigmoid"):
    """gvp.py:89-116 (vector gating variant). s [b,n], v [b,vi,3]."""
    Vh = torch.einsum("bvc,vh->bhc", v, sd[p + ".Wh"])
    Vu = torch.einsum("bhc,hu->buc", Vh, sd[p + ".Wu"])
    sh = norm_no_nan(Vh)
    f = F.silu(F.linear(torch.cat((s, sh), dim=1), sd[p + ".to_feats_out.0.weight"], sd[p + ".to_feats_out.0.bias"]))
    gate = F.linear(f, sd[p + ".scalar_to_vector_gates.weight"], sd[p + ".scalar_to_vector_gates.bias"])
    if vec_act == "sigmoid":
        gate = torch.sigmoid(gate)
    return f, gate.unsqueeze(-1) * Vu


def gvp_layernorm(sd, p: str, s: torch.Tensor, v: torch.Tensor, eps: float = 1e-5):
    """gvp.py:159-166."""
    s = F.layer_norm(s, (s.shape[-1],), sd[p + ".feat_norm.weight"], sd[p + ".feat_norm.bias"], 1e-5)
    vn = norm_no_nan(v, keepdims=True, sqrt=False)
    vn = torch.sqrt(torch.mean(vn, dim=-2, keepdim=True) + eps) + eps
    return s, v / vn


def encoder(sd, p: str, feats: torch.Tensor, t_node: torch.Tensor) -> torch.Tensor:
    """dynamics_gvp.py:107-117,143-151: LayerNorm(SiLU(Linear([feats, t])))."""
    z = torch.cat([feats, t_node.view(-1, 1)], dim=1)
    z = F.silu(F.linear(z, sd[p + ".0.weight"], sd[p + ".0.bias"]))
    return F.layer_norm(z, (z.shape[-1],), sd[p + ".2.weight"], sd[p + ".2.bias"], 1e-5)


def edge_messages(sd, p: str, h_src, v_src, x_src, x_dst, n_gvps: int = 3):
    """gvp.py:472-480 (edge features) + :540-551 (message) for one edge type; inputs are per-edge rows."""
    x_diff = x_src - x_dst
    dij = norm_no_nan(x_diff, keepdims=True) + 1e-8
    x_diff = x_diff / dij
    d = rbf(dij.squeeze(1))
    v = torch.cat([x_diff.unsqueeze(1), v_src], dim=1)
    s = torch.cat([h_src, d], dim=1)
    for i in range(n_gvps):
        s, v = gvp(sd, f"{p}.{i}", s, v)
    return s, v


def mean_aggregate(msg: torch.Tensor, dst: torch.Tensor, n_dst: int) -> torch.Tensor:
    """DGL fn.mean: sum over in-edges / in_degree.clamp(min=1); zero for isolated nodes."""
    acc = torch.zeros((n_dst,) + tuple(msg.shape[1:]), dtype=msg.dtype)
    acc.index_add_(0, dst, msg)
    deg = torch.zeros(n_dst, dtype=msg.dtype)
    deg.index_add_(0, dst, torch.ones(dst.shape[0], dtype=msg.dtype))
    return acc / deg.clamp(min=1).view((n_dst,) + (1,) * (msg.dim() - 1))


def conv_layer(sd, p: str, feats: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
               edges: Dict[str, Tuple[torch.Tensor, torch.Tensor]], n_message_gvps=3, n_update_gvps=2,
               return_messages: bool = False, message_norm="mean", norm0: Optional[Dict[str, torch.Tensor]] = None):
    """GVPMultiEdgeConv.forward, gvp.py:459-538, eval mode (dropout = identity).  message_norm='mean' (dev.yml): mean over
    the in-edges per edge type; a positive number: SUM over the in-edges (gvp.py:386-389) divided by it (:512-517); 0: SUM
    divided by the per-node value norm0[ntype] (`degree_norms`, gvp.py:504-507).

    feats[ntype] = (h [N,128], x [N,3], v [N,16,3]); edges[etype] = (src, dst).
    """
    agg_s = {nt: None for nt in ("pharm", "prot")}
    agg_v = {nt: None for nt in ("pharm", "prot")}
    for (snt, et, dnt) in ETYPES:
        src, dst = edges[et]
        if src.numel() == 0:
            continue
        hs, xs, vs = feats[snt]
        xd = feats[dnt][1]
        ms, mv = edge_messages(sd, f"{p}.edge_message_fns.{snt}_{et}_{dnt}", hs[src], vs[src], xs[src], xd[dst],
                               n_message_gvps)
        n_dst = feats[dnt][0].shape[0]
        if message_norm == "mean":
            a_s = mean_aggregate(ms, dst, n_dst)
            a_v = mean_aggregate(mv, dst, n_dst)
        elif message_norm == 0:
            a_s = torch.zeros((n_dst,) + tuple(ms.shape[1:]), dtype=ms.dtype).index_add_(0, dst, ms) / norm0[dnt].view(-1, 1)
            a_v = torch.zeros((n_dst,) + tuple(mv.shape[1:]), dtype=mv.dtype).index_add_(0, dst, mv) / norm0[dnt].view(-1, 1, 1)
        else:
            a_s = torch.zeros((n_dst,) + tuple(ms.shape[1:]), dtype=ms.dtype).index_add_(0, dst, ms) / float(message_norm)
            a_v = torch.zeros((n_dst,) + tuple(mv.shape[1:]), dtype=mv.dtype).index_add_(0, dst, mv) / float(message_norm)
        agg_s[dnt] = a_s if agg_s[dnt] is None else agg_s[dnt] + a_s
        agg_v[dnt] = a_v if agg_v[dnt] is None else agg_v[dnt] + a_v
    out = {}
    for nt in ("pharm", "prot"):
        h, x, v = feats[nt]
        h = h + agg_s[nt]
        v = v + agg_v[nt]
        h, v = gvp_layernorm(sd, f"{p}.message_layer_norms.{nt}", h, v)
        rs, rv = h, v
        for i in range(n_update_gvps):
            rs, rv = gvp(sd, f"{p}.node_update_fns.{nt}.{i}", rs, rv)
        h = h + rs
        v = v + rv
        h, v = gvp_layernorm(sd, f"{p}.update_layer_norms.{nt}", h, v)
        out[nt] = (h, x, v)
    if return_messages:
        return out, agg_s, agg_v
    return out


def noise_head(sd, p: str, h: torch.Tensor, v: torch.Tensor, n_gvps: int = 4):
    """NoisePredictionBlock.forward, dynamics_gvp.py:37-42 (last GVP: 1 vector out, identity gate act)."""
    for i in range(n_gvps):
        h, v = gvp(sd, f"{p}.gvps.{i}", h, v, vec_act="identity" if i == n_gvps - 1 else "sigmoid")
    return F.linear(h, sd[p + ".to_scalar_output.weight"], sd[p + ".to_scalar_output.bias"]), v.squeeze(1)


# --------------------------------------------------------------------------- denoiser + sampler
class FlatBatch:
    """Same information as the reference's batched DGL heterograph, as flat tensors."""

    def __init__(self, prot_x, prot_h, prot_ptr, pharm_ptr, pp_src, pp_dst):
        self.prot_x = prot_x.clone()
        self.prot_h = prot_h
        self.prot_ptr = prot_ptr.long()
        self.pharm_ptr = pharm_ptr.long()
        self.pp = (pp_src.long(), pp_dst.long())
        self.prot_b = batch_index(self.prot_ptr)
        self.pharm_b = batch_index(self.pharm_ptr)
        self.pharm_x = None
        self.pharm_h = None

    @property
    def n_graphs(self):
        return self.prot_ptr.numel() - 1


def build_batch(pockets: List[Tuple[torch.Tensor, torch.Tensor]], sizes: List[List[int]], pp_cutoff: float = 3.5):
    """generate_pharmacophores.py:323-334 / pharmacodiff.py:538-556: one graph per (pocket, sample);
    the pocket's pp radius graph (protein_pharm_dataset.py:234-236) is replicated per copy."""
    px, ph, pptr, fptr, es, ed = [], [], [0], [0], [], []
    for (pos, onehot), szs in zip(pockets, sizes):
        n = pos.shape[0]
        s, d = radius_edges(pos, torch.tensor([0, n]), pp_cutoff, 100)
        for nf in szs:
            off = pptr[-1]
            px.append(pos)
            ph.append(onehot)
            es.append(s + off)
            ed.append(d + off)
            pptr.append(off + n)
            fptr.append(fptr[-1] + int(nf))
    return FlatBatch(torch.cat(px), torch.cat(ph), torch.tensor(pptr), torch.tensor(fptr), torch.cat(es), torch.cat(ed))


def radius_bipartite_edges(prot_x, prot_ptr, pharm_x, pharm_ptr, r: float, max_num_neighbors: int = 100):
    """radius(x=pharm, y=prot, r, max_num_neighbors) per graph (dynamics_gvp.py:211): for every protein atom (query) the
    pharmacophore nodes of its graph with squared distance < r*r, ascending index, at most max_num_neighbors.
    Returns (prot_idx, pharm_idx) int64."""
    qs, cs = [], []
    r2 = torch.tensor(float(r), dtype=torch.float32) * torch.tensor(float(r), dtype=torch.float32)
    for g in range(prot_ptr.numel() - 1):
        a, b = int(prot_ptr[g]), int(prot_ptr[g + 1])
        qa, qb = int(pharm_ptr[g]), int(pharm_ptr[g + 1])
        if b == a or qb == qa:
            continue
        hit = _sqdist(prot_x[a:b], pharm_x[qa:qb]) < r2
        hit = hit & (torch.cumsum(hit.to(torch.int64), dim=1) <= max_num_neighbors)
        pi, fi = torch.nonzero(hit, as_tuple=True)
        qs.append(pi + a)
        cs.append(fi + qa)
    if not qs:
        z = torch.zeros(0, dtype=torch.int64)
        return z, z.clone()
    return torch.cat(qs), torch.cat(cs)


def dynamic_edges(b: FlatBatch, ff_cutoff: float = 9.0, pf_k: int = 5, ff_k: int = 0, pf_cutoff: float = 8.0):
    """dynamics_gvp.py:187-215 (ff_k=0 -> radius, ff_k>0 -> kNN graph; pf_k>0 -> kNN, pf_k=0 -> radius(pharm, prot, r_pf);
    fp = reverse of pf)."""
    if ff_k > 0:
        ff_src, ff_dst = knn_graph_edges(b.pharm_x, b.pharm_ptr, ff_k)
    else:
        ff_src, ff_dst = radius_edges(b.pharm_x, b.pharm_ptr, ff_cutoff, 200)
    if pf_k == 0:
        c, q = radius_bipartite_edges(b.prot_x, b.prot_ptr, b.pharm_x, b.pharm_ptr, pf_cutoff, 100)
        return {"ff": (ff_src, ff_dst), "pf": (c, q), "fp": (q, c), "pp": b.pp}
    q, c = knn_edges(b.prot_x, b.prot_ptr, b.pharm_x, b.pharm_ptr, pf_k)
    return {"ff": (ff_src, ff_dst), "pf": (c, q), "fp": (q, c), "pp": b.pp}


def degree_norms(b: FlatBatch, edges, pf_k: int) -> Dict[str, torch.Tensor]:
    """message_norm = 0 (gvp.py:504-507): per graph, (edges of every type into the node type) / (nodes of the type) + 1,
    broadcast to the nodes.  The per-graph edge counts are the ones add_pharm_edges records (dynamics_gvp.py:219-221): ff by
    the graph of the edge's first index (true counts); pp from the batched dataset graph (true counts); pf -- and fp, which
    copies it -- by `prot_batch_idx[pf_idxs[0]]`: with radius edges (pf_k = 0) pf_idxs[0] holds protein atoms (true counts),
    with kNN edges it holds PHARMACOPHORE node indices, so the edges of pharmacophore node i are counted for the graph that
    owns PROTEIN ATOM i.  That is what the reference computes; it is restated here, not corrected."""
    B = b.n_graphs
    cnt = lambda idx, owner: torch.bincount(owner[idx], minlength=B)
    e_ff = cnt(edges["ff"][0], b.pharm_b)
    e_pp = cnt(edges["pp"][1], b.prot_b)
    if pf_k > 0:
        e_pf = cnt(edges["pf"][1], b.prot_b)        # pharm node index looked up in the PROTEIN batch index (see above)
    else:
        e_pf = cnt(edges["pf"][0], b.prot_b)
    n_f = (b.pharm_ptr[1:] - b.pharm_ptr[:-1]).long()
    n_p = (b.prot_ptr[1:] - b.prot_ptr[:-1]).long()
    return {"pharm": ((e_ff + e_pf) / n_f + 1)[b.pharm_b], "prot": ((e_pf + e_pp) / n_p + 1)[b.prot_b]}


def denoiser(sd, b: FlatBatch, t: torch.Tensor, cfg: dict, prefix: str = "dynamics", trace: Optional[dict] = None):
    """PharmRecDynamicsGVP.forward, dynamics_gvp.py:131-185 -> (eps_h [Nf,6], eps_x [Nf,3])."""
    vs = cfg.get("vector_size", 16)
    h_f = encoder(sd, f"{prefix}.pharm_encoder", b.pharm_h, t[b.pharm_b])
    h_p = encoder(sd, f"{prefix}.prot_encoder", b.prot_h, t[b.prot_b])
    feats = {"pharm": (h_f, b.pharm_x, torch.zeros(h_f.shape[0], vs, 3)),
             "prot": (h_p, b.prot_x, torch.zeros(h_p.shape[0], vs, 3))}
    edges = dynamic_edges(b, cfg["graph_cutoffs"]["ff"], cfg.get("pf_k", 5), cfg.get("ff_k", 0),
                          cfg["graph_cutoffs"].get("pf", 8.0))
    mn = cfg.get("message_norm", "mean")
    norm0 = degree_norms(b, edges, cfg.get("pf_k", 5)) if (not isinstance(mn, str) and mn == 0) else None
    if trace is not None:
        trace["edges"] = edges
        trace["enc"] = {k: v[0] for k, v in feats.items()}
        trace["norm0"] = norm0
    for li in range(cfg.get("n_convs", 2)):
        feats = conv_layer(sd, f"{prefix}.noise_predictor.conv_layers.{li}", feats, edges,
                           cfg.get("n_message_gvps", 3), cfg.get("n_update_gvps", 2), message_norm=mn, norm0=norm0)
        if trace is not None:
            trace[f"conv{li}"] = {k: (v[0], v[2]) for k, v in feats.items()}
    return noise_head(sd, f"{prefix}.noise_predictor.noise_predictor", feats["pharm"][0], feats["pharm"][2],
                      cfg.get("n_noise_gvps", 4))


def segment_mean(x: torch.Tensor, seg: torch.Tensor, n_seg: int) -> torch.Tensor:
    """dgl.readout_nodes(op='mean')."""
    acc = torch.zeros((n_seg, x.shape[1]), dtype=x.dtype)
    acc.index_add_(0, seg, x)
    cnt = torch.bincount(seg, minlength=n_seg).to(x.dtype).view(-1, 1)
    return acc / cnt


def remove_pharm_com(b: FlatBatch):
    """pharmacodiff.py:88-108 with com='pharmacophore': shifts pharm x_t AND prot x_0."""
    com = segment_mean(b.pharm_x, b.pharm_b, b.n_graphs)
    b.pharm_x = b.pharm_x - com[b.pharm_b]
    b.prot_x = b.prot_x - com[b.prot_b]


def reverse_step(sd, b: FlatBatch, s_int: int, T: int, gamma: torch.Tensor, cfg: dict, noise_x, noise_h,
                 endpoint_feat: bool = False, endpoint_coord: bool = False):
    """sample_p_zs_given_zt, pharmacodiff.py:380-431 (eps parameterisation, or the endpoint branch :413-418 per part)."""
    B = b.n_graphs
    s_arr = torch.full((B,), s_int).float() / T
    t_arr = torch.full((B,), s_int + 1).float() / T
    g_s, g_t = gamma_at(gamma, s_arr, T), gamma_at(gamma, t_arr, T)
    alpha_ts, var_terms, sigma_q = posterior_coefficients(g_s, g_t)
    eps_h, eps_x = denoiser(sd, b, t_arr, cfg)
    fb = b.pharm_b
    c1, c2 = endpoint_coefficients(g_s, g_t)
    c1, c2 = c1[fb].view(-1, 1), c2[fb].view(-1, 1)
    if endpoint_coord:
        mu_x = c1 * b.pharm_x + c2 * eps_x
    else:
        mu_x = b.pharm_x / alpha_ts[fb].view(-1, 1) - var_terms[fb].view(-1, 1) * eps_x
    if endpoint_feat:
        mu_h = c1 * b.pharm_h + c2 * eps_h
    else:
        mu_h = b.pharm_h / alpha_ts[fb].view(-1, 1) - var_terms[fb].view(-1, 1) * eps_h
    b.pharm_x = mu_x + sigma_q[fb].view(-1, 1) * noise_x
    b.pharm_h = mu_h + sigma_q[fb].view(-1, 1) * noise_h
    remove_pharm_com(b)
    return eps_h, eps_x


def sample(sd, b: FlatBatch, noise: torch.Tensor, T: int, gamma: torch.Tensor, cfg: dict,
           init_pharm_com: Optional[torch.Tensor] = None, norm_const: float = 1.0, record=None, steps=None,
           endpoint_feat: bool = False, endpoint_coord: bool = False):
    """sample_given_receptor, pharmacodiff.py:433-514.  noise [T+1, Nf, 9]: row 0 = initial z_T
    (x cols 0:3, h cols 3:9), row 1+i = the i-th loop iteration (s = T-1-i), x drawn before h."""
    init_prot_com = segment_mean(b.prot_x, b.prot_b, b.n_graphs)
    if init_pharm_com is None:
        init_pharm_com = init_prot_com
    b.prot_x = b.prot_x - init_pharm_com[b.prot_b]
    b.pharm_x = noise[0, :, 0:3].clone()
    b.pharm_h = noise[0, :, 3:9].clone()
    if record is not None:
        record.append((b.pharm_x.clone(), b.pharm_h.clone(), b.prot_x.clone()))
    n_steps = T if steps is None else steps
    for i, s in enumerate(reversed(range(T))):
        if i >= n_steps:
            break
        reverse_step(sd, b, s, T, gamma, cfg, noise[1 + i, :, 0:3], noise[1 + i, :, 3:9], endpoint_feat, endpoint_coord)
        if record is not None:
            record.append((b.pharm_x.clone(), b.pharm_h.clone(), b.prot_x.clone()))
    # final frame restore, pharmacodiff.py:480-488
    com = segment_mean(b.prot_x, b.prot_b, b.n_graphs)
    x0 = b.pharm_x - com[b.pharm_b] + init_prot_com[b.pharm_b]
    prot = b.prot_x - com[b.prot_b] + init_prot_com[b.prot_b]
    h0 = b.pharm_h * norm_const
    return x0, h0, h0.argmax(dim=1), prot


def visual_frames(b: FlatBatch, record, init_prot_com: torch.Tensor):
    """get_pos_feat_for_visual, pharmacodiff.py:360-378, for every recorded state (x_t, h_t, prot x_0): the
    pharmacophore is moved back to the input frame by init_prot_com - current protein COM; h_t is stored as is
    (`unnormalize` touches h_0 only, :84-86).  -> (pos [F, Nf, 3], feat [F, Nf, nh])."""
    pos, feat = [], []
    for x_t, h_t, prot in record:
        prot_com = segment_mean(prot, b.prot_b, b.n_graphs)
        pos.append(x_t + (init_prot_com - prot_com)[b.pharm_b])
        feat.append(h_t.clone())
    return torch.stack(pos), torch.stack(feat)


def sample_multi(sd, pockets: List[Tuple[torch.Tensor, torch.Tensor]], n_pharms: List[List[int]], noise: torch.Tensor,
                 T: int, gamma: torch.Tensor, cfg: dict, max_batch_size: int = 32,
                 init_pharm_com: Optional[torch.Tensor] = None, norm_const: float = 1.0, frames: bool = False,
                 steps=None):
    """PharmacophoreDiff.sample, pharmacodiff.py:516-578: one graph per (pocket, requested size), pocket-major; chunks of
    max_batch_size consecutive graphs go through sample_given_receptor with init_pharm_com indexed by the graph's
    pocket (default: each pocket's plain mean position, :531-535); results are regrouped per pocket.
    noise [T+1, Nf_total, 9] is consumed by columns in the flattened graph order.
    -> list (per pocket) of lists of dicts {x, h, type, pos_frames?, feat_frames?}."""
    if init_pharm_com is None:
        init_pharm_com = torch.stack([pos.mean(dim=0) for pos, _ in pockets])
    flat = [(p, int(n)) for p, szs in enumerate(n_pharms) for n in szs]
    results, col = [], 0
    for start in range(0, len(flat), max_batch_size):
        chunk = flat[start:start + max_batch_size]
        # a chunk as its own batch: pockets in chunk order, one graph per entry
        b = build_batch([pockets[p] for p, _ in chunk], [[n] for _, n in chunk], cfg["graph_cutoffs"]["pp"])
        nf = int(b.pharm_ptr[-1])
        coms = init_pharm_com[[p for p, _ in chunk]]
        init_prot_com = segment_mean(b.prot_x, b.prot_b, b.n_graphs)
        rec = [] if frames else None
        x0, h0, typ, _ = sample(sd, b, noise[:, col:col + nf], T, gamma, cfg, init_pharm_com=coms,
                                norm_const=norm_const, record=rec, steps=steps)
        fr = visual_frames(b, rec, init_prot_com) if frames else None
        for gi in range(b.n_graphs):
            sl = slice(int(b.pharm_ptr[gi]), int(b.pharm_ptr[gi + 1]))
            out = {"x": x0[sl], "h": h0[sl], "type": typ[sl]}
            if fr is not None:
                out["pos_frames"], out["feat_frames"] = fr[0][:, sl], fr[1][:, sl]
            results.append(out)
        col += nf
    per_pocket, end = [], 0
    for szs in n_pharms:
        per_pocket.append(results[end:end + len(szs)])
        end += len(szs)
    return per_pocket


def forward_loss(sd, b: FlatBatch, x0: torch.Tensor, h0: torch.Tensor, t_int: torch.Tensor, eps_x: torch.Tensor,
                 eps_h: torch.Tensor, T: int, gamma: torch.Tensor, cfg: dict, norm_const: float = 1.0,
                 weighted_loss: bool = False, phase: str = "train", endpoint_feat: bool = False,
                 endpoint_coord: bool = False, remove_com: bool = True):
    """PharmacophoreDiff.forward, pharmacodiff.py:162-243: normalise h_0 (:81-83), remove the pharmacophore COM of x_0 from
    the complex (:178), noise with (t, eps) (:110-127), remove the COM of x_t (unless remove_com is off, :123-125),
    predict, and form the two losses and the four metrics -- MSE on eps (dev.yml), or, per part, the endpoint
    parameterisation (:204-216): cross entropy on the predicted h_0 / MSE on the predicted x_0 (+ the COM of x_t).
    t_int [B] in [0, T) and eps are injected."""
    fb = b.pharm_b
    h0 = h0 / norm_const
    com0 = segment_mean(x0, fb, b.n_graphs)
    x0 = x0 - com0[fb]
    b.prot_x = b.prot_x - com0[b.prot_b]
    t = t_int.float() / T
    g_t = gamma_at(gamma, t, T)
    alpha_t = torch.sqrt(torch.sigmoid(-g_t))[fb].view(-1, 1)
    sigma_t = torch.sqrt(torch.sigmoid(g_t))[fb].view(-1, 1)
    b.pharm_x = alpha_t * x0 + sigma_t * eps_x
    b.pharm_h = alpha_t * h0 + sigma_t * eps_h
    com_t = None
    if remove_com:
        com_t = segment_mean(b.pharm_x, fb, b.n_graphs)[fb]
        remove_pharm_com(b)
    h_dyn, x_dyn = denoiser(sd, b, t, cfg)
    if endpoint_feat:
        h0_pred = h_dyn
        h_loss = F.cross_entropy(h0_pred, h0.argmax(dim=1), reduction="none")
    else:
        h_loss = (eps_h - h_dyn).square().sum(dim=1)
        h0_pred = (b.pharm_h - sigma_t * h_dyn) / alpha_t
    if endpoint_coord:
        x0_pred = x_dyn + com_t if remove_com else x_dyn
        x_loss = (x0_pred - x0).square().sum(dim=1)
    else:
        x_loss = (eps_x - x_dyn).square().sum(dim=1)
        x0_pred = (b.pharm_x - sigma_t * x_dyn) / alpha_t
    w_metric = 1 - t[fb]
    w_loss = w_metric if weighted_loss else torch.ones_like(w_metric)
    losses = {phase + " pos loss": (x_loss * w_loss).sum() / eps_x.numel(),
              phase + " feat loss": (h_loss * w_loss).sum() / eps_h.numel()}
    sq = (x0_pred - x0).square().sum(dim=1)
    hit = (h0_pred.argmax(dim=1) == h0.argmax(dim=1)).float()
    metrics = {phase + " position error": sq.mean(), phase + " weighted position error": (w_metric * sq).mean(),
               phase + " accuracy": hit.mean(), phase + " weighted accuracy": (w_metric * hit).mean()}
    return losses, metrics


def nominal_edge_evals(b: FlatBatch, n_ff: int, pf_k: int = 5, n_convs: int = 2) -> int:
    """Edges pushed through the message MLP by one denoiser call (metric M2)."""
    nf = int(b.pharm_ptr[-1])
    return n_convs * (b.pp[0].numel() + 2 * pf_k * nf + n_ff)
