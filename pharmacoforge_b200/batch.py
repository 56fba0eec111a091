"""Flat device-resident batch of protein-pharmacophore graphs.

Carries the information of the reference's batched `dgl.DGLHeteroGraph` (node types prot / pharm, edge types
pp / pf / ff / fp; protein_pharm_dataset.py:210-266, unorganized_utils.py:28-95) as flat tensors:

  * graph g owns the contiguous protein nodes [prot_ptr[g], prot_ptr[g+1]) and pharmacophore nodes
    [pharm_ptr[g], pharm_ptr[g+1]) -- the layout `dgl.batch` produces and `get_batch_idxs` exposes;
  * the static pp radius graph is a destination-sorted CSR over all protein nodes of the batch, with its
    tile plan, built once per batch on the GPU (K1);
  * the per-step pharmacophore edges (ff radius, pf kNN, fp reverse) live in fixed-capacity buffers that
    K2 refills every reverse-diffusion step, so no structure is ever mutated on the host.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops


@dataclass
class Pocket:
    """One receptor pocket: the reference's `ref_graph` restricted to what the hot path reads."""
    prot_x: torch.Tensor   # [N,3] float32
    prot_h: torch.Tensor   # [N,F] float32 (element one-hot)

    @staticmethod
    def from_numpy(pos: np.ndarray, onehot: np.ndarray) -> "Pocket":
        return Pocket(torch.from_numpy(np.ascontiguousarray(pos, dtype=np.float32)),
                      torch.from_numpy(np.ascontiguousarray(onehot, dtype=np.float32)))

    @staticmethod
    def from_dgl(g) -> "Pocket":
        """Adapter for a reference pocket graph (requires dgl; same node data names as the reference)."""
        return Pocket(g.nodes["prot"].data["x_0"].detach().float().cpu().contiguous(),
                      g.nodes["prot"].data["h_0"].detach().float().cpu().contiguous())


    def to_dgl(self, graph_cutoffs: dict):
        """The reference's pocket graph for this pocket (io.pocket_to_dgl); requires dgl."""
        from .io import pocket_to_dgl
        return pocket_to_dgl(self, graph_cutoffs)


MAX_PHARM_PER_GRAPH = 128   # PF_MAX_PHARM_PER_GRAPH of include/pharmacoforge_b200.h


def _chunk_graphs(weights: np.ndarray, target: int) -> np.ndarray:
    """Group consecutive graphs into planner chunks of roughly `target` edge rows (boundaries in graphs)."""
    bounds = [0]
    acc = 0
    for g, w in enumerate(weights):
        acc += int(w)
        if acc >= target:
            bounds.append(g + 1)
            acc = 0
    if bounds[-1] != len(weights):
        bounds.append(len(weights))
    return np.asarray(bounds, dtype=np.int64)


def on_batch_device(fn):
    """Method decorator for `(self, g: GraphBatch, ...)` entry points: the kernels launch on the CURRENT device, so the
    batch's device is entered for the duration of the call (a batch on cuda:1 works whatever the caller's current
    device is; ops.py rejects tensors that do not live on the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, g, *args, **kwargs):
        with torch.cuda.device(g.device):
            return fn(self, g, *args, **kwargs)
    return wrapper


class GraphBatch:
    """See module docstring.  Build with `GraphBatch.from_pockets`."""

    def __init__(self):
        self.device = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_pockets(cls, pockets: Sequence[Pocket], sizes: Sequence[Sequence[int]], device, pp_cutoff: float = 3.5,
                     pf_k: int = 5, ff_max_nbrs: int = 200, pp_max_nbrs: int = 100,
                     graph_range: Optional[range] = None, tile_rows: int = 128, pf_max_nbrs: int = 100) -> "GraphBatch":
        """One graph per (pocket, requested pharmacophore size), pocket-major, exactly the order
        `PharmacophoreDiff.sample` flattens them in (pharmacodiff.py:538-544).  `graph_range` restricts the batch
        to a slice of that flattened list (max_batch_size chunking / multi-GPU sharding)."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("GraphBatch lives on a CUDA device; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):   # the kernels launch on the current device
            return cls._from_pockets(pockets, sizes, dev, pp_cutoff, pf_k, ff_max_nbrs, pp_max_nbrs, graph_range,
                                     tile_rows, pf_max_nbrs)

    @classmethod
    def _from_pockets(cls, pockets, sizes, dev, pp_cutoff, pf_k, ff_max_nbrs, pp_max_nbrs, graph_range, tile_rows,
                      pf_max_nbrs=100):
        self = cls()
        self.device = dev
        self.pf_k, self.ff_max_nbrs, self.pf_max_nbrs = int(pf_k), int(ff_max_nbrs), int(pf_max_nbrs)
        if self.pf_k < 0 or self.pf_max_nbrs < 1:
            raise ValueError("pf_k >= 0 and pf_max_nbrs >= 1")
        tile_rows = int(os.environ.get("PF_TILE_ROWS", tile_rows))   # A/B switch: 64 = fp32 FFMA kernels
        if tile_rows not in (64, 128):
            raise ValueError("tile_rows: 128 (tcgen05 kernels) or 64 (fp32 FFMA kernels)")
        self.tile_rows = int(tile_rows)
        graph_pocket, graph_nf = [], []
        for p, szs in enumerate(sizes):
            for nf in szs:
                graph_pocket.append(p)
                graph_nf.append(int(nf))
        if graph_range is not None:
            graph_pocket = graph_pocket[graph_range.start:graph_range.stop]
            graph_nf = graph_nf[graph_range.start:graph_range.stop]
        B = len(graph_pocket)
        if any(nf < 1 or nf > MAX_PHARM_PER_GRAPH for nf in graph_nf):
            raise ValueError(f"pharmacophore sizes must be in 1..{MAX_PHARM_PER_GRAPH} (PF_MAX_PHARM_PER_GRAPH: the per-step "
                             "graph kernel keeps a graph's centres in shared memory)")
        used = sorted(set(graph_pocket))
        remap = {p: i for i, p in enumerate(used)}
        graph_pocket = np.asarray([remap[p] for p in graph_pocket], dtype=np.int64)
        graph_nf = np.asarray(graph_nf, dtype=np.int64)
        pk_n = np.asarray([pockets[p].prot_x.shape[0] for p in used], dtype=np.int64)
        pk_off = np.concatenate([[0], np.cumsum(pk_n)])
        graph_np = pk_n[graph_pocket] if B else np.zeros(0, dtype=np.int64)
        prot_ptr = np.concatenate([[0], np.cumsum(graph_np)]).astype(np.int32)
        pharm_ptr = np.concatenate([[0], np.cumsum(graph_nf)]).astype(np.int32)
        self.n_graphs, self.n_prot, self.n_pharm = B, int(prot_ptr[-1]), int(pharm_ptr[-1])
        self.graph_pocket = graph_pocket
        self.pocket_ids = used
        self.prot_ptr_host, self.pharm_ptr_host = prot_ptr, pharm_ptr

        # ---- host -> device: only the distinct pockets travel; replication happens on the GPU
        pk_x = torch.cat([pockets[p].prot_x for p in used]).float().contiguous()
        pk_h = torch.cat([pockets[p].prot_h for p in used]).float().contiguous()
        self.n_prot_feats = pk_h.shape[1]
        meta = torch.from_numpy(np.concatenate([prot_ptr, pharm_ptr, pk_off[graph_pocket].astype(np.int32),
                                                pk_off.astype(np.int32)]))
        self.h2d_bytes = pk_x.numel() * 4 + pk_h.numel() * 4 + meta.numel() * 4
        pk_x = pk_x.pin_memory().to(dev, non_blocking=True)
        pk_h = pk_h.pin_memory().to(dev, non_blocking=True)
        meta = meta.pin_memory().to(dev, non_blocking=True)
        self.prot_ptr = meta[:B + 1].contiguous()
        self.pharm_ptr = meta[B + 1:2 * B + 2].contiguous()
        g_pk_off = meta[2 * B + 2:3 * B + 2].long()
        pk_ptr = meta[3 * B + 2:].contiguous()            # node ranges of the distinct pockets
        counts = (self.prot_ptr[1:] - self.prot_ptr[:-1]).long()
        node_graph = torch.repeat_interleave(torch.arange(B, device=dev), counts, output_size=self.n_prot)
        src_row = torch.arange(self.n_prot, device=dev) - self.prot_ptr[:-1].long()[node_graph] + g_pk_off[node_graph]
        self.prot_x = pk_x[src_row].contiguous()
        self.prot_feats = pk_h[src_row].contiguous()
        self.prot_x0 = self.prot_x.clone()      # input frame, kept for the final restore / re-use of the batch
        self.pharm_x = torch.zeros(self.n_pharm, 3, device=dev)
        self.pharm_h = None                     # allocated by the sampler (feature width is a model property)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)

        # ---- K1: static pp radius graph.  Built ONCE PER DISTINCT POCKET with a cell list, then replicated per graph
        # with node offsets on the device -- what protein_pharm_dataset.py:234-236 (one radius_graph per pocket) +
        # copy_graph / dgl.batch (unorganized_utils.py:28-50) do.  PF_K1=brute: the all-pairs kernel over the
        # replicated batch (same CSR bit for bit; kept as the A/B reference of the tests).
        self._k1_inputs = (pk_x, pk_ptr, g_pk_off, counts, float(pp_cutoff), int(pp_max_nbrs))
        self.pp_rowptr, self.pp_cnt, self.pp_col = self.build_pp_graph()
        self.pp_start = self.pp_rowptr[:-1]
        self.n_pp_edges = int(self.pp_col.numel())
        self.pp_tiles = torch.empty(2 * max(self.n_prot + B, 1), dtype=torch.int32, device=dev)
        self.pp_n_tiles = torch.zeros(1, dtype=torch.int32, device=dev)
        # tiles in graph order (PF_PP_PLAN=atomic: the unordered planner, the A/B switch of the L2-reuse measurement)
        plan = ops.plan_tiles if os.environ.get("PF_PP_PLAN", "ordered") == "atomic" else ops.plan_tiles_ordered
        plan(self.pp_cnt, self.prot_ptr, False, self.tile_rows, self.pp_tiles, self.pp_n_tiles, self.status)
        self.pp_num_tiles = int(self.pp_n_tiles.item())
        self.check_status()   # e.g. PF_DEV_DEGREE_OVERFLOW: a pp in-degree above the tile capacity, reported now
        self.pp_tiles = self.pp_tiles[:2 * max(self.pp_num_tiles, 1)].clone()

        # ---- static description of the dynamic edge buffers (K2 refills them every step)
        k = self.pf_k
        nf = graph_nf
        ff_base = np.concatenate([[0], np.cumsum(nf * np.maximum(nf - 1, 0))])
        local = np.arange(self.n_pharm) - np.repeat(pharm_ptr[:-1].astype(np.int64), nf)
        ff_start = (np.repeat(ff_base[:-1], nf) + local * np.repeat(np.maximum(nf - 1, 0), nf)).astype(np.int32)
        self.ff_capacity = int(ff_base[-1])
        # planner chunks of ~PF_PLAN_CHUNK edge rows (whole graphs): a tile never spans two chunks, so a chunk's last tile is
        # partly filled -- 1,024 rows keep that to ~1 tile in 9 (256, the value of the one-thread-per-chunk planner: 1 in 3)
        gb = _chunk_graphs(np.maximum(nf * np.maximum(nf - 1, 0), k * nf), int(os.environ.get("PF_PLAN_CHUNK", 1024)))
        pharm_chunk = pharm_ptr[gb].astype(np.int32)
        fp_chunk = (k * pharm_ptr[gb].astype(np.int64)).astype(np.int32)
        small = torch.from_numpy(np.concatenate([ff_start, (k * np.arange(self.n_pharm)).astype(np.int32), pharm_chunk,
                                                 fp_chunk])).to(dev)
        n = self.n_pharm
        self.ff_start = small[:n].contiguous()
        self.pf_start = small[n:2 * n].contiguous()
        self.pharm_chunk_ptr = small[2 * n:2 * n + len(gb)].contiguous()
        self.fp_chunk_ptr = small[2 * n + len(gb):].contiguous()
        self.n_chunks = len(gb) - 1
        i32 = dict(dtype=torch.int32, device=dev)
        self.ff_cnt = torch.zeros(max(n, 1), **i32)
        self.ff_col = torch.zeros(max(self.ff_capacity, 1), **i32)
        self.pf_cnt = torch.zeros(max(n, 1), **i32)
        self.pf_col = torch.zeros(max(k * n, 1), **i32)
        self.fp_seg_dst = torch.zeros(max(k * n, 1), **i32)
        self.fp_seg_start = torch.zeros(max(k * n, 1), **i32)
        self.fp_seg_cnt = torch.zeros(max(k * n, 1), **i32)
        self.fp_col = torch.zeros(max(k * n, 1), **i32)
        self.n_fp_chunks = self.n_chunks
        self.dyn_max_tiles = k * n + self.n_chunks + 1
        if k == 0:
            self._radius_buffers(graph_nf, graph_np, local)
        self.ff_tiles = torch.zeros(2 * self.dyn_max_tiles, **i32)
        self.pf_tiles = torch.zeros(2 * self.dyn_max_tiles, **i32)
        self.fp_tiles = torch.zeros(2 * self.dyn_max_tiles, **i32)
        self.dyn_n_tiles = torch.zeros(3, **i32)
        return self

    def _radius_buffers(self, nf, np_g, local):
        """pf_k == 0 (dynamics_gvp.py:210-216): pf / fp edges from radius(pharm, prot, r_pf, pf_max_nbrs per protein atom),
        refilled every step by pf_dyn_graph_radius.  A pharmacophore node can collect every atom of its pocket, more than an
        edge tile holds, so its pf segment has capacity n_prot(graph) and is cut into ceil(n_prot(graph) / tile_rows)
        sub-segment slots (the edge kernels write one mean per slot, pf_combine_subsegments folds them); fp has one
        segment per protein atom with capacity min(nf, pf_max_nbrs).  Replaces the kNN-mode pf / fp buffers."""
        dev, n, B, rows = self.device, self.n_pharm, self.n_graphs, self.tile_rows
        i32 = dict(dtype=torch.int32, device=dev)
        n_sub_g = -(-np_g // rows)
        pf_base = np.concatenate([[0], np.cumsum(nf * np_g)])
        sub_base = np.concatenate([[0], np.cumsum(nf * n_sub_g)])
        fp_cap = np.minimum(nf, self.pf_max_nbrs)
        fp_base = np.concatenate([[0], np.cumsum(np_g * fp_cap)])
        if max(int(pf_base[-1]), int(fp_base[-1])) >= 2 ** 31:
            raise ValueError("pf_k == 0: the radius pf / fp edge buffers of this batch exceed int32 indexing; use smaller batches")
        pf_start = np.repeat(pf_base[:-1], nf) + local * np.repeat(np_g, nf)
        sub_ptr = np.concatenate([np.repeat(sub_base[:-1], nf) + local * np.repeat(n_sub_g, nf), sub_base[-1:]])
        small = torch.from_numpy(np.concatenate([pf_start, sub_ptr, fp_base[:-1], sub_base]).astype(np.int32)).to(dev)
        self.pf_start = small[:n].contiguous()
        self.pf_sub_ptr = small[n:2 * n + 1].contiguous()
        self.fp_base = small[2 * n + 1:2 * n + 1 + B].contiguous()
        self.pf_sub_chunk_ptr = small[2 * n + 1 + B:].contiguous()       # planner chunks: the sub-segment slots of one graph
        self.n_pf_sub = int(sub_base[-1])
        self.pf_col = torch.zeros(max(int(pf_base[-1]), 1), **i32)
        self.pf_sub_start = torch.zeros(max(self.n_pf_sub, 1), **i32)
        self.pf_sub_cnt = torch.zeros(max(self.n_pf_sub, 1), **i32)
        self.pf_sub_x = torch.zeros(max(self.n_pf_sub, 1), 3, dtype=torch.float32, device=dev)
        self.fp_seg_start = torch.zeros(max(self.n_prot, 1), **i32)
        self.fp_seg_cnt = torch.zeros(max(self.n_prot, 1), **i32)
        self.fp_col = torch.zeros(max(int(fp_base[-1]), 1), **i32)
        self.fp_chunk_ptr = self.prot_ptr                                 # fp planner chunks: the atoms of one graph
        self.n_fp_chunks = B
        # tiles: ff <= one per node; pf <= one per sub-segment slot; fp: two consecutive tiles of a chunk hold more than
        # tile_rows edges or tile_rows segments between them
        fp_tiles = int(np.minimum(np_g, 2 * (-(-(np_g * fp_cap) // rows)) + -(-np_g // rows) + 2).sum())
        self.dyn_max_tiles = max(n + self.n_chunks + 1, self.n_pf_sub + 1, fp_tiles + 1)

    def set_pharmacophores(self, x_0: torch.Tensor, h_0: torch.Tensor) -> "GraphBatch":
        """Ground-truth pharmacophores of a training / validation batch: `g.nodes['pharm'].data['x_0' / 'h_0']` of the
        reference (protein_pharm_dataset.py:238-243), node-major in graph order.  Only `PharmacophoreDiff.forward`
        reads them; sampling ignores them."""
        if x_0.shape != (self.n_pharm, 3) or h_0.shape[0] != self.n_pharm:
            raise ValueError(f"expected x_0 [{self.n_pharm}, 3] and h_0 [{self.n_pharm}, F]")
        self.pharm_x0 = x_0.to(self.device, torch.float32).contiguous()
        self.pharm_h0 = h_0.to(self.device, torch.float32).contiguous()
        return self

    def build_pp_graph(self):
        """K1 (see from_pockets): -> (rowptr [N+1], cnt [N], col [E]) of the batched static pp graph.  Re-runnable: bench.py
        times it for the K1 roofline line."""
        pk_x, pk_ptr, g_pk_off, counts, pp_cutoff, pp_max_nbrs = self._k1_inputs
        if os.environ.get("PF_K1", "cell") == "brute" or self.n_graphs == 0:
            return ops.radius_csr(self.prot_x, self.prot_ptr, pp_cutoff, pp_max_nbrs)
        pk_rowptr, pk_cnt, pk_col = ops.cell_radius_csr(pk_x, pk_ptr, pp_cutoff, pp_max_nbrs)
        self.pk_csr = (pk_rowptr, pk_cnt, pk_col)      # kept: the shared-pocket mode runs the first layer's pp messages on it
        node0 = g_pk_off.to(torch.int32)
        graph_edges = (pk_rowptr[g_pk_off + counts] - pk_rowptr[g_pk_off]).to(torch.int32)
        edge0 = ops.exclusive_scan(graph_edges)
        n_edges = int(edge0[-1].item())
        return ops.replicate_csr(pk_rowptr, pk_col, node0, self.prot_ptr, edge0, self.n_prot, n_edges)

    def share_arrays(self):
        """Static arrays of the opt-in shared-pocket mode (dynamics.share_pocket_messages, csrc/pf_share.cu), or None when
        the batch does not qualify (protein features not one-hot / built with the all-pairs K1): the distinct pockets'
        coordinates, pp CSR and tile plan, the seed-table row of every distinct node, the first distinct node of every
        graph's pocket, and the identity inputs of the per-(graph, type) encoder table."""
        cached = getattr(self, "_share_arrays", False)
        if cached is not False:
            return cached
        self._share_arrays = None
        if self.seed_arrays() is None or getattr(self, "pk_csr", None) is None or self.n_graphs == 0:
            return None
        dev = self.device
        pk_x, pk_ptr, g_pk_off, counts, _, _ = self._k1_inputs
        pk_rowptr, pk_cnt, pk_col = self.pk_csr
        n_d, F, B = int(pk_x.shape[0]), self.n_prot_feats, self.n_graphs
        tiles = torch.empty(2 * max(n_d + pk_ptr.numel(), 1), dtype=torch.int32, device=dev)
        n_tiles = torch.zeros(1, dtype=torch.int32, device=dev)
        ops.plan_tiles_ordered(pk_cnt, pk_ptr, False, self.tile_rows, tiles, n_tiles, self.status)
        n = int(n_tiles.item())
        # first graph of every distinct pocket (graphs are pocket-major): its (graph, type) rows seed the pocket's messages
        first_graph = np.full(len(self.pocket_ids), -1, dtype=np.int64)
        for gi in range(B - 1, -1, -1):
            first_graph[self.graph_pocket[gi]] = gi
        pk_sizes = (pk_ptr[1:] - pk_ptr[:-1]).long()
        node_pocket = torch.repeat_interleave(torch.arange(pk_sizes.numel(), device=dev), pk_sizes, output_size=n_d)
        # one-hot type of every distinct node: read it back from the replicated features of the pocket's first graph
        fg = torch.from_numpy(first_graph).to(dev)
        local = torch.arange(n_d, device=dev) - pk_ptr[:-1].long()[node_pocket]
        rep_node = self.prot_ptr[:-1].long()[fg[node_pocket]] + local
        types = self.prot_feats[rep_node].argmax(dim=1)
        pk_seed_row = (fg[node_pocket] * F + types).to(torch.int32).contiguous()
        self._share_arrays = dict(
            pk_x=pk_x.contiguous(), pk_start=pk_rowptr[:-1].contiguous(), pk_cnt=pk_cnt, pk_col=pk_col,
            pk_tiles=tiles[:2 * max(n, 1)].clone(), pk_n_tiles=n_tiles, pk_max_tiles=n, n_distinct=n_d,
            pk_seed_row=pk_seed_row, pk_node0=g_pk_off.to(torch.int32).contiguous(),
            **self.enc_arrays())
        return self._share_arrays

    def enc_arrays(self):
        """Inputs of the per-(graph, atom type) encoder table (one-hot protein features): identity blocks as the encoder's
        feature rows, F rows per graph, every table row its own representative."""
        cached = getattr(self, "_enc_arrays", None)
        if cached is None:
            dev, F, B = self.device, self.n_prot_feats, self.n_graphs
            cached = self._enc_arrays = dict(
                enc_feats=torch.eye(F, device=dev).repeat(B, 1).contiguous(),
                enc_ptr=(F * torch.arange(B + 1, device=dev)).to(torch.int32),
                enc_rep=torch.arange(B * F, device=dev, dtype=torch.int32))
        return cached

    def seed_arrays(self):
        """(seed_row [n_prot] int32, seed_rep [n_graphs * F] int32) for the first-layer seeding of the pp messages, or
        None when the protein features are not one-hot.  The encoder output of a one-hot row depends only on (graph, atom
        type) -- SURVEY.md §8 a4 -- so protein node n reads table row graph(n) * F + type(n), and row r is computed from
        the representative node seed_rep[r] (-1: no atom of that type in that graph)."""
        cached = getattr(self, "_seed_arrays", False)
        if cached is not False:
            return cached
        f = self.prot_feats
        F = f.shape[1]
        onehot = bool(((f == 0) | (f == 1)).all().item()) and bool((f.sum(dim=1) == 1).all().item())
        if not onehot:
            self._seed_arrays = None
            return None
        row = self.batch_idxs()["prot"] * F + f.argmax(dim=1)
        rep = torch.full((self.n_graphs * F,), self.n_prot, dtype=torch.int64, device=self.device)
        rep.scatter_reduce_(0, row, torch.arange(self.n_prot, device=self.device), reduce="amin")
        rep[rep == self.n_prot] = -1
        self._seed_arrays = (row.to(torch.int32).contiguous(), rep.to(torch.int32).contiguous())
        return self._seed_arrays

    # ------------------------------------------------------------------ reference-style accessors
    def batch_idxs(self):
        """unorganized_utils.get_batch_idxs: graph index of every node, per node type."""
        B = self.n_graphs
        ar = torch.arange(B, device=self.device)
        return {"prot": torch.repeat_interleave(ar, (self.prot_ptr[1:] - self.prot_ptr[:-1]).long(),
                                                output_size=self.n_prot),
                "pharm": torch.repeat_interleave(ar, (self.pharm_ptr[1:] - self.pharm_ptr[:-1]).long(),
                                                 output_size=self.n_pharm)}

    @property
    def batch_size(self):
        return self.n_graphs

    def check_status(self):
        from . import _lib
        _lib.check_dev_status(int(self.status.item()) & 0xFFFFFFFF)

    def pp_edges(self):
        """(src, dst) int64 of the static pp graph, in (dst, src) order -- for tests."""
        dst = torch.repeat_interleave(torch.arange(self.n_prot, device=self.device), self.pp_cnt.long())
        return self.pp_col.long(), dst

    def dynamic_edges(self):
        """Current ff / pf / fp edge lists as (src, dst) int64 -- for tests (host sync)."""
        k = self.pf_k
        n = self.n_pharm
        ar = torch.arange(n, device=self.device)
        out = {}
        cnt = self.ff_cnt[:n].long()
        dst = torch.repeat_interleave(ar, cnt)
        within = torch.arange(dst.numel(), device=self.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
        out["ff"] = (self.ff_col[self.ff_start.long()[dst] + within].long(), dst)
        cnt = self.pf_cnt[:n].long()
        dst = torch.repeat_interleave(ar, cnt)
        within = torch.arange(dst.numel(), device=self.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
        if k == 0:    # radius mode: whole pf segments at pf_start, one fp segment per protein atom
            out["pf"] = (self.pf_col[self.pf_start.long()[dst] + within].long(), dst)
            scnt = self.fp_seg_cnt[:self.n_prot].long()
            seg = torch.repeat_interleave(torch.arange(self.n_prot, device=self.device), scnt)
            within = torch.arange(seg.numel(), device=self.device) - torch.repeat_interleave(torch.cumsum(scnt, 0) - scnt, scnt)
            out["fp"] = (self.fp_col[self.fp_seg_start.long()[seg] + within].long(), seg)
            return out
        out["pf"] = (self.pf_col[k * dst + within].long(), dst)
        scnt = self.fp_seg_cnt[:k * n].long()
        seg = torch.repeat_interleave(torch.arange(k * n, device=self.device), scnt)
        within = torch.arange(seg.numel(), device=self.device) - torch.repeat_interleave(torch.cumsum(scnt, 0) - scnt, scnt)
        out["fp"] = (self.fp_col[self.fp_seg_start.long()[seg] + within].long(), self.fp_seg_dst.long()[seg])
        return out
