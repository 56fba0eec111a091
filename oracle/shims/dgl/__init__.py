"""Pure-PyTorch stand-in for the slice of DGL that PharmacoForge's hot path touches.

TEST INFRASTRUCTURE ONLY.  DGL is an un-vendored third-party dependency of the
reference (install_things.sh:6, unpinned) and is not installable in this
container.  This shim restates the documented semantics of exactly the calls
the reference makes, so that the reference's own
`pharmacoforge/models/*.py` can be imported UNMODIFIED from /root/reference and
executed on CPU to generate golden vectors (oracle/make_golden.py).

Semantics restated (DGL docs):
  * heterograph / batch / unbatch: graph g of a batch owns a contiguous node-id
    range per node type and a contiguous edge-id range per edge type.
  * apply_edges(fn.u_sub_v(a, b, out)): out[e] = src[a][u_e] - dst[b][v_e].
  * apply_edges(udf): udf(EdgeBatch) with .src/.dst/.data/.canonical_etype.
  * multi_update_all({etype: (copy_e, mean|sum)}, 'sum'): per-etype reduce onto
    dst nodes (mean = sum / in_degree.clamp(min=1)), etypes with zero edges are
    skipped, results sharing a dst type are summed.
  * add_edges appends and DROPS batch info (hence the reference's save/restore
    at dynamics_gvp.py:189,224-225); remove_edges deletes by edge id.
  * local_scope reverts feature writes, not structure.
  * readout_nodes(op='mean'): per-graph segment mean.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Tuple

import torch

from . import function  # noqa: F401
from . import data  # noqa: F401
from . import dataloading  # noqa: F401

__version__ = "0.0-shim"


class _Frame(dict):
    pass


class _TypedView:
    """g.nodes[...] / g.edges[...] -> object with a `.data` frame."""

    def __init__(self, frame):
        self.data = frame


class _NodeView:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        return _TypedView(self._g._ndata[ntype])


class _EdgeView:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, etype):
        return _TypedView(self._g._edata[self._g.to_canonical_etype(etype)])

    def __call__(self, form="uv", etype=None):
        ce = self._g.to_canonical_etype(etype)
        u, v = self._g._edges[ce]
        if form == "uv":
            return u, v
        if form == "eid":
            return torch.arange(u.shape[0], device=u.device)
        if form == "all":
            return u, v, torch.arange(u.shape[0], device=u.device)
        raise ValueError(form)


class EdgeBatch:
    def __init__(self, g, cetype):
        u, v = g._edges[cetype]
        s, _, d = cetype
        self._cetype = cetype
        self.src = {k: t[u] for k, t in g._ndata[s].items()}
        self.dst = {k: t[v] for k, t in g._ndata[d].items()}
        self.data = g._edata[cetype]

    @property
    def canonical_etype(self):
        return self._cetype


class DGLHeteroGraph:
    def __init__(self, edges: Dict[Tuple[str, str, str], Tuple[torch.Tensor, torch.Tensor]],
                 num_nodes: Dict[str, int], device=None):
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self._num_nodes = {k: int(v) for k, v in num_nodes.items()}
        self._ntypes = sorted(self._num_nodes)
        self._cetypes = sorted(edges, key=lambda c: c[1])
        self._edges = {}
        for ce in self._cetypes:
            u, v = edges[ce]
            u = torch.as_tensor(u, dtype=torch.int64, device=self._device).reshape(-1)
            v = torch.as_tensor(v, dtype=torch.int64, device=self._device).reshape(-1)
            self._edges[ce] = (u, v)
        self._ndata = {nt: _Frame() for nt in self._ntypes}
        self._edata = {ce: _Frame() for ce in self._cetypes}
        self._batch_num_nodes = None
        self._batch_num_edges = None

    # ---- structure ----
    @property
    def ntypes(self):
        return list(self._ntypes)

    @property
    def canonical_etypes(self):
        return list(self._cetypes)

    @property
    def etypes(self):
        return [c[1] for c in self._cetypes]

    @property
    def device(self):
        return self._device

    @property
    def nodes(self):
        return _NodeView(self)

    @property
    def edges(self):
        return _EdgeView(self)

    def to_canonical_etype(self, etype):
        if isinstance(etype, tuple):
            return etype
        for ce in self._cetypes:
            if ce[1] == etype:
                return ce
        raise KeyError(etype)

    def num_nodes(self, ntype=None):
        return self._num_nodes[ntype]

    def num_edges(self, etype=None):
        return int(self._edges[self.to_canonical_etype(etype)][0].shape[0])

    @property
    def batch_size(self):
        if self._batch_num_nodes is None:
            return 1
        return int(next(iter(self._batch_num_nodes.values())).shape[0])

    def batch_num_nodes(self, ntype=None):
        if self._batch_num_nodes is None:
            return torch.tensor([self._num_nodes[ntype]], dtype=torch.int64, device=self._device)
        return self._batch_num_nodes[ntype]

    def batch_num_edges(self, etype=None):
        ce = self.to_canonical_etype(etype)
        if self._batch_num_edges is None:
            return torch.tensor([self.num_edges(ce)], dtype=torch.int64, device=self._device)
        return self._batch_num_edges[ce]

    def set_batch_num_nodes(self, val):
        self._batch_num_nodes = {k: torch.as_tensor(v, dtype=torch.int64) for k, v in val.items()}

    def set_batch_num_edges(self, val):
        self._batch_num_edges = {self.to_canonical_etype(k): torch.as_tensor(v, dtype=torch.int64)
                                 for k, v in val.items()}

    def add_edges(self, u, v, etype=None):
        ce = self.to_canonical_etype(etype)
        u0, v0 = self._edges[ce]
        u = torch.as_tensor(u, dtype=torch.int64, device=self._device).reshape(-1)
        v = torch.as_tensor(v, dtype=torch.int64, device=self._device).reshape(-1)
        self._edges[ce] = (torch.cat([u0, u]), torch.cat([v0, v]))
        n_new = u.shape[0]
        for k, t in list(self._edata[ce].items()):
            pad = torch.zeros((n_new,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            self._edata[ce][k] = torch.cat([t, pad])
        self._batch_num_nodes = None
        self._batch_num_edges = None

    def remove_edges(self, eids, etype=None):
        ce = self.to_canonical_etype(etype)
        u0, v0 = self._edges[ce]
        keep = torch.ones(u0.shape[0], dtype=torch.bool, device=u0.device)
        keep[torch.as_tensor(eids, dtype=torch.int64)] = False
        self._edges[ce] = (u0[keep], v0[keep])
        for k, t in list(self._edata[ce].items()):
            self._edata[ce][k] = t[keep]
        self._batch_num_nodes = None
        self._batch_num_edges = None

    def to(self, device):
        device = torch.device(device)
        g = DGLHeteroGraph({ce: (u.to(device), v.to(device)) for ce, (u, v) in self._edges.items()},
                           self._num_nodes, device=device)
        for nt in self._ntypes:
            for k, t in self._ndata[nt].items():
                g._ndata[nt][k] = t.to(device)
        for ce in self._cetypes:
            for k, t in self._edata[ce].items():
                g._edata[ce][k] = t.to(device)
        if self._batch_num_nodes is not None:
            g._batch_num_nodes = {k: v.to(device) for k, v in self._batch_num_nodes.items()}
        if self._batch_num_edges is not None:
            g._batch_num_edges = {k: v.to(device) for k, v in self._batch_num_edges.items()}
        return g

    @contextlib.contextmanager
    def local_scope(self):
        nd = {nt: _Frame(f) for nt, f in self._ndata.items()}
        ed = {ce: _Frame(f) for ce, f in self._edata.items()}
        try:
            yield
        finally:
            # feature writes are reverted; structural edits (add/remove_edges) persist
            self._ndata = nd
            for ce in self._cetypes:
                n_e = self._edges[ce][0].shape[0]
                self._edata[ce] = _Frame({k: t for k, t in ed[ce].items() if t.shape[0] == n_e})

    # ---- message passing ----
    def apply_edges(self, func, etype=None):
        ce = self.to_canonical_etype(etype)
        if isinstance(func, function._BinaryMsg):
            u, v = self._edges[ce]
            a = self._ndata[ce[0]][func.lhs][u]
            b = self._ndata[ce[2]][func.rhs][v]
            self._edata[ce][func.out] = func.op(a, b)
            return
        out = func(EdgeBatch(self, ce))
        for k, t in out.items():
            self._edata[ce][k] = t

    def _reduce_etype(self, ce, msg, red):
        u, v = self._edges[ce]
        m = self._edata[ce][msg.field]
        n_dst = self._num_nodes[ce[2]]
        acc = torch.zeros((n_dst,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        acc.index_add_(0, v, m)
        if red.kind == "mean":
            deg = torch.zeros(n_dst, dtype=m.dtype, device=m.device)
            deg.index_add_(0, v, torch.ones_like(v, dtype=m.dtype))
            deg = deg.clamp(min=1).reshape((n_dst,) + (1,) * (m.dim() - 1))
            acc = acc / deg
        return acc

    def update_all(self, msg, red, etype=None):
        ce = self.to_canonical_etype(etype)
        self._ndata[ce[2]][red.out] = self._reduce_etype(ce, msg, red)

    def multi_update_all(self, etype_dict, cross_reducer):
        assert cross_reducer == "sum"
        per_dst: Dict[str, List[torch.Tensor]] = {}
        out_name = None
        for etype, (msg, red) in etype_dict.items():
            ce = self.to_canonical_etype(etype)
            if self.num_edges(ce) == 0:
                continue
            per_dst.setdefault(ce[2], []).append(self._reduce_etype(ce, msg, red))
            out_name = red.out
        for nt, parts in per_dst.items():
            self._ndata[nt][out_name] = torch.stack(parts, dim=0).sum(dim=0)


DGLGraph = DGLHeteroGraph


def heterograph(data_dict, num_nodes_dict=None, idtype=None, device=None):
    edges = {}
    for ce, (u, v) in data_dict.items():
        edges[ce] = (torch.as_tensor(u, dtype=torch.int64), torch.as_tensor(v, dtype=torch.int64))
    if device is None:
        for u, _ in edges.values():
            if isinstance(u, torch.Tensor) and u.numel():
                device = u.device
                break
    return DGLHeteroGraph(edges, num_nodes_dict, device=device)


def batch(graphs):
    g0 = graphs[0]
    dev = g0.device
    offs = {nt: 0 for nt in g0.ntypes}
    eu = {ce: [] for ce in g0.canonical_etypes}
    ev = {ce: [] for ce in g0.canonical_etypes}
    bnn = {nt: [] for nt in g0.ntypes}
    bne = {ce: [] for ce in g0.canonical_etypes}
    for g in graphs:
        for ce in g0.canonical_etypes:
            u, v = g._edges[ce]
            eu[ce].append(u + offs[ce[0]])
            ev[ce].append(v + offs[ce[2]])
            bne[ce].append(g.batch_num_edges(ce))
        for nt in g0.ntypes:
            bnn[nt].append(g.batch_num_nodes(nt))
            offs[nt] += g.num_nodes(nt)
    out = DGLHeteroGraph({ce: (torch.cat(eu[ce]), torch.cat(ev[ce])) for ce in eu}, offs, device=dev)
    for nt in g0.ntypes:
        for k in g0._ndata[nt]:
            out._ndata[nt][k] = torch.cat([g._ndata[nt][k] for g in graphs], dim=0)
    for ce in g0.canonical_etypes:
        for k in g0._edata[ce]:
            out._edata[ce][k] = torch.cat([g._edata[ce][k] for g in graphs], dim=0)
    out._batch_num_nodes = {nt: torch.cat(bnn[nt]).to(dev) for nt in bnn}
    out._batch_num_edges = {ce: torch.cat(bne[ce]).to(dev) for ce in bne}
    return out


def unbatch(g):
    B = g.batch_size
    res = []
    noff = {nt: 0 for nt in g.ntypes}
    eoff = {ce: 0 for ce in g.canonical_etypes}
    for b in range(B):
        nn = {nt: int(g.batch_num_nodes(nt)[b]) for nt in g.ntypes}
        edges = {}
        for ce in g.canonical_etypes:
            ne = int(g.batch_num_edges(ce)[b])
            u, v = g._edges[ce]
            s = slice(eoff[ce], eoff[ce] + ne)
            edges[ce] = (u[s] - noff[ce[0]], v[s] - noff[ce[2]])
        gi = DGLHeteroGraph(edges, nn, device=g.device)
        for nt in g.ntypes:
            s = slice(noff[nt], noff[nt] + nn[nt])
            for k, t in g._ndata[nt].items():
                gi._ndata[nt][k] = t[s]
        for ce in g.canonical_etypes:
            ne = int(g.batch_num_edges(ce)[b])
            s = slice(eoff[ce], eoff[ce] + ne)
            for k, t in g._edata[ce].items():
                gi._edata[ce][k] = t[s]
            eoff[ce] += ne
        for nt in g.ntypes:
            noff[nt] += nn[nt]
        res.append(gi)
    return res


def readout_nodes(g, feat, weight=None, *, op="sum", ntype=None):
    x = g._ndata[ntype][feat]
    counts = g.batch_num_nodes(ntype)
    B = counts.shape[0]
    seg = torch.arange(B, device=x.device).repeat_interleave(counts)
    acc = torch.zeros((B,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    acc.index_add_(0, seg, x)
    if op == "sum":
        return acc
    if op == "mean":
        return acc / counts.to(x.dtype).reshape((B,) + (1,) * (x.dim() - 1))
    raise ValueError(op)
