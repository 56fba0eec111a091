"""torch.library custom ops over the C ABI (include/pharmacoforge_b200.h).

Every op enqueues hand-written sm_100a kernels on torch's current CUDA stream through ctypes; tensors are
passed as raw device pointers.  There is no CPU implementation and no dispatch on architecture: calling an
op with a non-CUDA tensor raises.  Ops that write into caller-provided buffers declare them in
`mutates_args`; shape/dtype-only fakes are registered so the ops can be traced.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

_L = _lib.load()
NS = "pharmacoforge"


import threading

_call = threading.local()   # devices of the tensors marshalled for the library call being assembled


def _p(t: Optional[torch.Tensor], dtype=None):
    """Device address of a tensor argument (checked: CUDA, contiguous, dtype) and a note of its device for `_s`."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.PfError("pharmacoforge ops need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise _lib.PfError("pharmacoforge ops need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise _lib.PfError(f"expected {dtype}, got {t.dtype}")
    try:
        _call.devs.add(t.device.index)
    except AttributeError:
        _call.devs = {t.device.index}
    return t.data_ptr()      # ctypes converts the int for the c_void_p parameter


def _f(t):
    return _p(t, torch.float32)


def _i(t):
    return _p(t, torch.int32)


def _s():
    """Stream argument of a library call; always the LAST argument, so every tensor of the call has been marshalled.
    All of them must live on one device and that device must be the current one (the kernels launch on the current
    device): the public entry points (`GraphBatch.from_pockets`, `PharmRecDynamicsGVP.forward`,
    `PharmacophoreDiff.sample_given_receptor / forward`) enter `torch.cuda.device(batch.device)` themselves."""
    devs = getattr(_call, "devs", None) or set()
    _call.devs = set()
    cur = torch.cuda.current_device()
    if len(devs) > 1:
        raise _lib.PfError(f"pharmacoforge ops need all tensors on one device, got cuda:{sorted(devs)}")
    if devs and next(iter(devs)) != cur:
        raise _lib.PfError(f"tensors live on cuda:{next(iter(devs))} but the current device is cuda:{cur}: wrap the "
                           "call in torch.cuda.device(...)")
    # raw handle of torch's current stream on the current device (torch.cuda.current_stream() builds a Stream object per
    # call: ~15 us, 550 times per training step)
    return torch._C._cuda_getCurrentRawStream(cur)


# ------------------------------------------------------------------------------------------------ graph
@torch.library.custom_op(f"{NS}::exclusive_scan", mutates_args=())
def exclusive_scan(x: torch.Tensor) -> torch.Tensor:
    n = x.numel()
    if n == 0:
        return torch.zeros(1, dtype=torch.int32, device=x.device)
    out = torch.empty(n + 1, dtype=torch.int32, device=x.device)
    ws_bytes = _L.pf_scan_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.int32, device=x.device)
    _lib.check(_L.pf_exclusive_scan_i32(_i(x), _i(out), n, _p(ws), ws.numel() * 4, _s()), "pf_exclusive_scan_i32")
    return out


@exclusive_scan.register_fake
def _(x):
    return x.new_empty(x.numel() + 1)


@torch.library.custom_op(f"{NS}::radius_count", mutates_args=())
def radius_count(x: torch.Tensor, seg_ptr: torch.Tensor, r: float, max_nbrs: int) -> torch.Tensor:
    deg = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    _lib.check(_L.pf_radius_count(_f(x), _i(seg_ptr), seg_ptr.numel() - 1, r, max_nbrs, _i(deg), _s()),
               "pf_radius_count")
    return deg


@radius_count.register_fake
def _(x, seg_ptr, r, max_nbrs):
    return x.new_empty(x.shape[0], dtype=torch.int32)


@torch.library.custom_op(f"{NS}::radius_fill", mutates_args=())
def radius_fill(x: torch.Tensor, seg_ptr: torch.Tensor, r: float, max_nbrs: int, rowptr: torch.Tensor,
                n_edges: int) -> torch.Tensor:
    col = torch.empty(max(n_edges, 1), dtype=torch.int32, device=x.device)
    _lib.check(_L.pf_radius_fill(_f(x), _i(seg_ptr), seg_ptr.numel() - 1, r, max_nbrs, _i(rowptr), _i(col), _s()),
               "pf_radius_fill")
    return col[:n_edges]


@radius_fill.register_fake
def _(x, seg_ptr, r, max_nbrs, rowptr, n_edges):
    return x.new_empty(n_edges, dtype=torch.int32)


def radius_csr(x: torch.Tensor, seg_ptr: torch.Tensor, r: float, max_nbrs: int):
    """K1: destination-sorted CSR of the radius graph within each segment -> (rowptr [N+1], deg [N], col [E])."""
    deg = radius_count(x, seg_ptr, r, max_nbrs)
    rowptr = exclusive_scan(deg)
    n_edges = int(rowptr[-1].item())  # one host sync per batch, at setup
    col = radius_fill(x, seg_ptr, r, max_nbrs, rowptr, n_edges)
    return rowptr, deg, col


@torch.library.custom_op(f"{NS}::cell_radius_csr", mutates_args=())
def cell_radius_csr(x: torch.Tensor, seg_ptr: torch.Tensor, r: float, max_nbrs: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """K1 as a cell list: destination-sorted CSR of the radius graph within each segment (pocket) ->
    (rowptr [N+1], deg [N], col [E]); bit-identical to `radius_csr`."""
    n, n_seg = x.shape[0], seg_ptr.numel() - 1
    ws_bytes = _L.pf_cell_radius_workspace_bytes(n, n_seg)
    ws = torch.empty(max(ws_bytes // 4, 1), dtype=torch.int32, device=x.device)
    deg = torch.empty(n, dtype=torch.int32, device=x.device)
    _lib.check(_L.pf_cell_radius_count(_f(x), _i(seg_ptr), n_seg, n, r, max_nbrs, _p(ws), ws.numel() * 4, _i(deg), _s()),
               "pf_cell_radius_count")
    rowptr = exclusive_scan(deg)
    n_edges = int(rowptr[-1].item())  # one host sync per batch, at setup
    col = torch.empty(max(n_edges, 1), dtype=torch.int32, device=x.device)
    _lib.check(_L.pf_cell_radius_fill(_f(x), _i(seg_ptr), n_seg, n, r, max_nbrs, _p(ws), ws.numel() * 4, _i(rowptr), _i(col),
                                      _s()), "pf_cell_radius_fill")
    return rowptr, deg, col[:n_edges]


@cell_radius_csr.register_fake
def _(x, seg_ptr, r, max_nbrs):
    return (x.new_empty(x.shape[0] + 1, dtype=torch.int32), x.new_empty(x.shape[0], dtype=torch.int32),
            x.new_empty(0, dtype=torch.int32))


@torch.library.custom_op(f"{NS}::replicate_csr", mutates_args=())
def replicate_csr(pk_rowptr: torch.Tensor, pk_col: torch.Tensor, pk_node0: torch.Tensor, prot_ptr: torch.Tensor,
                  edge0: torch.Tensor, n_nodes: int, n_edges: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """copy_graph + dgl.batch of the static pp graph: per-pocket CSR -> batched (rowptr [N+1], cnt [N], col [E])."""
    dev = pk_rowptr.device
    rowptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=dev)
    cnt = torch.empty(max(n_nodes, 1), dtype=torch.int32, device=dev)
    col = torch.empty(max(n_edges, 1), dtype=torch.int32, device=dev)
    _lib.check(_L.pf_replicate_csr(_i(pk_rowptr), _i(pk_col), _i(pk_node0), _i(prot_ptr), _i(edge0), prot_ptr.numel() - 1,
                                   _i(rowptr), _i(cnt), _i(col), _s()), "pf_replicate_csr")
    return rowptr, cnt[:n_nodes], col[:n_edges]


@replicate_csr.register_fake
def _(pk_rowptr, pk_col, pk_node0, prot_ptr, edge0, n_nodes, n_edges):
    return (pk_rowptr.new_empty(n_nodes + 1), pk_rowptr.new_empty(n_nodes), pk_rowptr.new_empty(n_edges))


@torch.library.custom_op(f"{NS}::dyn_graph",
                         mutates_args=("ff_cnt", "ff_col", "pf_cnt", "pf_col", "fp_seg_dst", "fp_seg_start",
                                       "fp_seg_cnt", "fp_col", "status"))
def dyn_graph(prot_x: torch.Tensor, prot_ptr: torch.Tensor, pharm_x: torch.Tensor, pharm_ptr: torch.Tensor,
              ff_r: float, ff_max_nbrs: int, pf_k: int, ff_start: torch.Tensor, ff_cnt: torch.Tensor,
              ff_col: torch.Tensor, pf_cnt: torch.Tensor, pf_col: torch.Tensor, fp_seg_dst: torch.Tensor,
              fp_seg_start: torch.Tensor, fp_seg_cnt: torch.Tensor, fp_col: torch.Tensor,
              status: torch.Tensor, ff_k: int = 0) -> None:
    """K2.  ff_k > 0: ff edges = knn_graph(pharm x_t, k = ff_k) (dynamics_gvp.py:194) instead of the radius graph."""
    _lib.check(_L.pf_dyn_graph_ffk(_f(prot_x), _i(prot_ptr), _f(pharm_x), _i(pharm_ptr), prot_ptr.numel() - 1, ff_r,
                                   ff_max_nbrs, ff_k, pf_k, _i(ff_start), _i(ff_cnt), _i(ff_col), _i(pf_cnt), _i(pf_col),
                                   _i(fp_seg_dst), _i(fp_seg_start), _i(fp_seg_cnt), _i(fp_col), _p(status), _s()),
               "pf_dyn_graph")


@torch.library.custom_op(f"{NS}::dyn_graph_radius",
                         mutates_args=("ff_cnt", "ff_col", "pf_cnt", "pf_col", "sub_start", "sub_cnt", "sub_x", "fp_seg_start",
                                       "fp_seg_cnt", "fp_col", "status"))
def dyn_graph_radius(prot_x: torch.Tensor, prot_ptr: torch.Tensor, pharm_x: torch.Tensor, pharm_ptr: torch.Tensor,
                     ff_r: float, ff_max_nbrs: int, ff_k: int, pf_r: float, pf_max_nbrs: int, sub_rows: int,
                     ff_start: torch.Tensor, ff_cnt: torch.Tensor, ff_col: torch.Tensor, pf_start: torch.Tensor,
                     sub_ptr: torch.Tensor, fp_base: torch.Tensor, pf_cnt: torch.Tensor, pf_col: torch.Tensor,
                     sub_start: torch.Tensor, sub_cnt: torch.Tensor, sub_x: torch.Tensor, fp_seg_start: torch.Tensor,
                     fp_seg_cnt: torch.Tensor,
                     fp_col: torch.Tensor, status: torch.Tensor) -> None:
    """K2 with pf_k == 0: pf / fp edges from radius(pharm, prot, r_pf, max per protein atom) (dynamics_gvp.py:210-216)."""
    _lib.check(_L.pf_dyn_graph_radius(_f(prot_x), _i(prot_ptr), _f(pharm_x), _i(pharm_ptr), prot_ptr.numel() - 1, ff_r,
                                      ff_max_nbrs, ff_k, pf_r, pf_max_nbrs, sub_rows, _i(ff_start), _i(ff_cnt), _i(ff_col),
                                      _i(pf_start), _i(sub_ptr), _i(fp_base), _i(pf_cnt), _i(pf_col), _i(sub_start),
                                      _i(sub_cnt), _f(sub_x), _i(fp_seg_start), _i(fp_seg_cnt), _i(fp_col), _p(status), _s()),
               "pf_dyn_graph_radius")


@torch.library.custom_op(f"{NS}::combine_subsegments", mutates_args=("agg_h", "agg_v"))
def combine_subsegments(sub_h: torch.Tensor, sub_v: torch.Tensor, sub_cnt: torch.Tensor, sub_ptr: torch.Tensor,
                        tot_cnt: torch.Tensor, inv_norm: float, agg_h: torch.Tensor, agg_v: torch.Tensor,
                        accumulate: bool, inv_norm_node: Optional[torch.Tensor] = None) -> None:
    """agg[d] (+)= sum of the sub-segment means of d weighted by count / total (inv_norm == 0) or count * inv_norm."""
    _lib.check(_L.pf_combine_subsegments(_f(sub_h), _f(sub_v), _i(sub_cnt), _i(sub_ptr), _i(tot_cnt), agg_h.shape[0],
                                         inv_norm, _f(inv_norm_node), _f(agg_h), _f(agg_v), int(accumulate), _s()),
               "pf_combine_subsegments")


@torch.library.custom_op(f"{NS}::plan_tiles", mutates_args=("tiles", "n_tiles", "status"))
def plan_tiles(seg_cnt: torch.Tensor, chunk_ptr: torch.Tensor, skip_empty: bool, tile_rows: int, tiles: torch.Tensor,
               n_tiles: torch.Tensor, status: torch.Tensor) -> None:
    _lib.check(_L.pf_zero_i32(_i(n_tiles), 1, _s()), "pf_zero_i32")
    _lib.check(_L.pf_plan_tiles(_i(seg_cnt), _i(chunk_ptr), chunk_ptr.numel() - 1, int(skip_empty), tile_rows,
                                _i(tiles), tiles.numel() // 2, _i(n_tiles), _p(status), _s()), "pf_plan_tiles")


@torch.library.custom_op(f"{NS}::plan_tiles_ordered", mutates_args=("tiles", "n_tiles", "status"))
def plan_tiles_ordered(seg_cnt: torch.Tensor, chunk_ptr: torch.Tensor, skip_empty: bool, tile_rows: int,
                       tiles: torch.Tensor, n_tiles: torch.Tensor, status: torch.Tensor) -> None:
    """plan_tiles with the tile list in chunk (graph) order: count per chunk, exclusive scan, fill (static plans)."""
    n_chunks = chunk_ptr.numel() - 1
    per_chunk = torch.empty(max(n_chunks, 1), dtype=torch.int32, device=seg_cnt.device)
    _lib.check(_L.pf_plan_tiles_count(_i(seg_cnt), _i(chunk_ptr), n_chunks, int(skip_empty), tile_rows, _i(per_chunk),
                                      _p(status), _s()), "pf_plan_tiles_count")
    off = exclusive_scan(per_chunk[:n_chunks])
    _lib.check(_L.pf_zero_i32(_i(n_tiles), 1, _s()), "pf_zero_i32")
    _lib.check(_L.pf_plan_tiles_fill(_i(seg_cnt), _i(chunk_ptr), n_chunks, int(skip_empty), tile_rows, _i(off), _i(tiles),
                                     tiles.numel() // 2, _i(n_tiles), _p(status), _s()), "pf_plan_tiles_fill")


# ------------------------------------------------------------------------------------------------ compute
@torch.library.custom_op(f"{NS}::encode", mutates_args=())
def encode(feats: torch.Tensor, node_ptr: torch.Tensor, t: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    out = torch.empty(feats.shape[0], 128, dtype=torch.float32, device=feats.device)
    _lib.check(_L.pf_encode(_f(feats), feats.shape[1], _i(node_ptr), node_ptr.numel() - 1, _f(t), _f(w), _f(out),
                            _s()), "pf_encode")
    return out


@encode.register_fake
def _(feats, node_ptr, t, w):
    return feats.new_empty(feats.shape[0], 128)


@torch.library.custom_op(f"{NS}::edge_conv", mutates_args=("agg_h", "agg_v"))
def edge_conv(src_h: torch.Tensor, src_v: Optional[torch.Tensor], src_x: torch.Tensor, dst_x: torch.Tensor,
              seg_start: torch.Tensor, seg_cnt: torch.Tensor, seg_dst: Optional[torch.Tensor], col: torch.Tensor,
              tiles: torch.Tensor, n_tiles: torch.Tensor, w: torch.Tensor, n_gvps: int, agg_h: torch.Tensor,
              agg_v: torch.Tensor, accumulate: bool) -> None:
    _lib.check(_L.pf_edge_conv(_f(src_h), _f(src_v), _f(src_x), _f(dst_x), _i(seg_start), _i(seg_cnt), _i(seg_dst),
                               _i(col), _i(tiles), _i(n_tiles), tiles.numel() // 2, _f(w), n_gvps, _f(agg_h),
                               _f(agg_v), int(accumulate), _s()), "pf_edge_conv")


@torch.library.custom_op(f"{NS}::edge_conv_tc", mutates_args=("agg_h", "agg_v"))
def edge_conv_tc(src_h: torch.Tensor, src_v: Optional[torch.Tensor], src_x: torch.Tensor, dst_x: torch.Tensor,
                 seg_start: torch.Tensor, seg_cnt: torch.Tensor, seg_dst: Optional[torch.Tensor], col: torch.Tensor,
                 tiles: torch.Tensor, n_tiles: torch.Tensor, wblob: torch.Tensor, agg_h: torch.Tensor,
                 agg_v: torch.Tensor, accumulate: bool, fp16: bool = False) -> None:
    """K3 on the tensor cores (tcgen05): tiles must be planned with tile_rows=128; wblob from pack_message_tc.
    fp16=True: the single-pass reduced-precision variant (pf_edge_conv_tc_f16)."""
    if wblob.dtype != torch.uint8 or wblob.numel() != _L.pf_tc_msg_blob_bytes():
        raise _lib.PfError("edge_conv_tc: wblob must be the uint8 image built by weights.pack_message_tc")
    _lib.check((_L.pf_edge_conv_tc_f16 if fp16 else _L.pf_edge_conv_tc)(_f(src_h), _f(src_v), _f(src_x), _f(dst_x), _i(seg_start), _i(seg_cnt), _i(seg_dst),
                                  _i(col), _i(tiles), _i(n_tiles), tiles.numel() // 2, _p(wblob), _f(agg_h),
                                  _f(agg_v), int(accumulate), _s()), "pf_edge_conv_tc")


@torch.library.custom_op(f"{NS}::edge_conv_tc_mapped", mutates_args=("agg_h", "agg_v"))
def edge_conv_tc_mapped(src_h: torch.Tensor, src_map: torch.Tensor, src_v: Optional[torch.Tensor], src_x: torch.Tensor,
                        dst_x: torch.Tensor, seg_start: torch.Tensor, seg_cnt: torch.Tensor, seg_dst: Optional[torch.Tensor],
                        col: torch.Tensor, tiles: torch.Tensor, n_tiles: torch.Tensor, wblob: torch.Tensor,
                        agg_h: torch.Tensor, agg_v: torch.Tensor, accumulate: bool, fp16: bool = False) -> None:
    """K3 (general tcgen05 kernel) whose source scalars come from a row table: node n reads row src_map[n] of src_h."""
    if wblob.dtype != torch.uint8 or wblob.numel() != _L.pf_tc_msg_blob_bytes():
        raise _lib.PfError("edge_conv_tc_mapped: wblob must be the uint8 image built by weights.pack_message_tc")
    _lib.check(_L.pf_edge_conv_tc_mapped(_f(src_h), _i(src_map), _f(src_v), _f(src_x), _f(dst_x), _i(seg_start), _i(seg_cnt),
                                         _i(seg_dst), _i(col), _i(tiles), _i(n_tiles), tiles.numel() // 2, _p(wblob),
                                         _f(agg_h), _f(agg_v), int(accumulate), int(fp16), _s()), "pf_edge_conv_tc_mapped")


@torch.library.custom_op(f"{NS}::seed_table", mutates_args=("table",))
def seed_table(h: torch.Tensor, rep_node: torch.Tensor, w_msg: torch.Tensor, table: torch.Tensor) -> None:
    """table[r] = k Wf0[:, 0:128] h[rep_node[r]] (rows with rep_node[r] < 0 untouched); w_msg = fp32 packed message chain."""
    _lib.check(_L.pf_seed_table(_f(h), _i(rep_node), rep_node.numel(), _f(w_msg), _f(table), _s()), "pf_seed_table")


@torch.library.custom_op(f"{NS}::edge_conv_tc_seeded", mutates_args=("agg_h", "agg_v"))
def edge_conv_tc_seeded(seed_row: torch.Tensor, table: torch.Tensor, src_x: torch.Tensor, dst_x: torch.Tensor,
                        seg_start: torch.Tensor, seg_cnt: torch.Tensor, seg_dst: Optional[torch.Tensor],
                        col: torch.Tensor, tiles: torch.Tensor, n_tiles: torch.Tensor, wblob: torch.Tensor,
                        agg_h: torch.Tensor, agg_v: torch.Tensor, accumulate: bool, fp16: bool = False) -> None:
    """K3 of the first conv layer with the per-node part of GVP 0 taken from `table` (see pf_seed_table)."""
    if wblob.dtype != torch.uint8 or wblob.numel() != _L.pf_tc_msg_blob_bytes():
        raise _lib.PfError("edge_conv_tc_seeded: wblob must be the uint8 image built by weights.pack_message_tc")
    _lib.check(_L.pf_edge_conv_tc_seeded(_i(seed_row), _f(table), _f(src_x), _f(dst_x), _i(seg_start), _i(seg_cnt),
                                         _i(seg_dst), _i(col), _i(tiles), _i(n_tiles), tiles.numel() // 2, _p(wblob),
                                         _f(agg_h), _f(agg_v), int(accumulate), int(fp16), _s()),
               "pf_edge_conv_tc_seeded")


@torch.library.custom_op(f"{NS}::node_update", mutates_args=("h_out", "v_out"))
def node_update(h_in: torch.Tensor, v_in: Optional[torch.Tensor], agg_h: torch.Tensor, agg_v: torch.Tensor,
                w: torch.Tensor, n_gvps: int, h_out: torch.Tensor, v_out: torch.Tensor) -> None:
    _lib.check(_L.pf_node_update(_f(h_in), _f(v_in), _f(agg_h), _f(agg_v), h_in.shape[0], _f(w), n_gvps, _f(h_out),
                                 _f(v_out), _s()), "pf_node_update")


@torch.library.custom_op(f"{NS}::node_update_tc", mutates_args=("h_out", "v_out"))
def node_update_tc(h_in: torch.Tensor, v_in: Optional[torch.Tensor], agg_h: torch.Tensor, agg_v: torch.Tensor,
                   wblob: torch.Tensor, h_out: torch.Tensor, v_out: torch.Tensor, fp16: bool = False) -> None:
    """K4 on the tensor cores (tcgen05); wblob from weights.pack_update_tc.  fp16=True: pf_node_update_tc_f16."""
    if wblob.dtype != torch.uint8 or wblob.numel() != _L.pf_tc_upd_blob_bytes():
        raise _lib.PfError("node_update_tc: wblob must be the uint8 image built by weights.pack_update_tc")
    _lib.check((_L.pf_node_update_tc_f16 if fp16 else _L.pf_node_update_tc)(_f(h_in), _f(v_in), _f(agg_h), _f(agg_v), h_in.shape[0], _p(wblob), _f(h_out),
                                    _f(v_out), _s()), "pf_node_update_tc")


@torch.library.custom_op(f"{NS}::node_update_tc_mapped", mutates_args=("h_out", "v_out"))
def node_update_tc_mapped(h_table: torch.Tensor, h_map: torch.Tensor, agg_h: torch.Tensor, agg_v: torch.Tensor,
                          wblob: torch.Tensor, h_out: torch.Tensor, v_out: torch.Tensor, fp16: bool = False) -> None:
    """K4 of the first conv layer (no input vectors) whose input scalars come from a row table: node n reads row
    h_map[n] of h_table; h_out / v_out are full per-node arrays."""
    if wblob.dtype != torch.uint8 or wblob.numel() != _L.pf_tc_upd_blob_bytes():
        raise _lib.PfError("node_update_tc_mapped: wblob must be the uint8 image built by weights.pack_update_tc")
    _lib.check(_L.pf_node_update_tc_mapped(_f(h_table), _i(h_map), None, _f(agg_h), _f(agg_v), h_map.numel(), _p(wblob),
                                           _f(h_out), _f(v_out), int(fp16), _s()), "pf_node_update_tc_mapped")


@torch.library.custom_op(f"{NS}::noise_head", mutates_args=())
def noise_head(h: torch.Tensor, v: torch.Tensor, w: torch.Tensor, n_gvps: int,
               n_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    n = h.shape[0]
    eps_h = torch.empty(n, n_out, dtype=torch.float32, device=h.device)
    eps_x = torch.empty(n, 3, dtype=torch.float32, device=h.device)
    _lib.check(_L.pf_noise_head(_f(h), _f(v), n, _f(w), n_gvps, n_out, _f(eps_h), _f(eps_x), _s()), "pf_noise_head")
    return eps_h, eps_x


@noise_head.register_fake
def _(h, v, w, n_gvps, n_out):
    return h.new_empty(h.shape[0], n_out), h.new_empty(h.shape[0], 3)


@torch.library.custom_op(f"{NS}::posterior_step", mutates_args=("pharm_x", "pharm_h", "prot_x"))
def posterior_step(pharm_x: torch.Tensor, pharm_h: torch.Tensor, eps_x: torch.Tensor, eps_h: torch.Tensor,
                   noise_x: torch.Tensor, noise_h: torch.Tensor, pharm_ptr: torch.Tensor, prot_x: torch.Tensor,
                   prot_ptr: torch.Tensor, alpha_ts: float, var_terms: float, sigma_q: float) -> None:
    _lib.check(_L.pf_posterior_step(_f(pharm_x), _f(pharm_h), pharm_h.shape[1], _f(eps_x), _f(eps_h), _f(noise_x),
                                    _f(noise_h), _i(pharm_ptr), _f(prot_x), _i(prot_ptr), prot_ptr.numel() - 1,
                                    alpha_ts, var_terms, sigma_q, _s()), "pf_posterior_step")


@torch.library.custom_op(f"{NS}::posterior_step_ep", mutates_args=("pharm_x", "pharm_h", "prot_x"))
def posterior_step_ep(pharm_x: torch.Tensor, pharm_h: torch.Tensor, pred_x: torch.Tensor, pred_h: torch.Tensor,
                      noise_x: torch.Tensor, noise_h: torch.Tensor, pharm_ptr: torch.Tensor, prot_x: torch.Tensor,
                      prot_ptr: torch.Tensor, alpha_ts: float, var_terms: float, sigma_q: float, ep_c1: float, ep_c2: float,
                      ep_mode: int) -> None:
    """sample_p_zs_given_zt with the endpoint parameterisation for the parts in ep_mode (1: coordinates, 2: features;
    pharmacodiff.py:413-418), injected noise."""
    _lib.check(_L.pf_posterior_step_ep(_f(pharm_x), _f(pharm_h), pharm_h.shape[1], _f(pred_x), _f(pred_h), _f(noise_x),
                                       _f(noise_h), None, 0, _i(pharm_ptr), _f(prot_x), _i(prot_ptr), prot_ptr.numel() - 1,
                                       alpha_ts, var_terms, sigma_q, ep_c1, ep_c2, ep_mode, _s()), "pf_posterior_step_ep")


@torch.library.custom_op(f"{NS}::philox_normal", mutates_args=("out",))
def philox_normal(out: torch.Tensor, seed: torch.Tensor, stream_id: int, step: int) -> None:
    """out <- N(0, 1) draws of (stream_id, step) under the int64 device seed (Philox4x32-10 + Box-Muller)."""
    if seed.dtype != torch.int64 or seed.numel() != 1:
        raise _lib.PfError("philox_normal: seed must be a 1-element int64 CUDA tensor")
    _lib.check(_L.pf_philox_normal(_f(out), out.numel(), _p(seed), stream_id, step, _s()), "pf_philox_normal")


@torch.library.custom_op(f"{NS}::segment_mean3", mutates_args=())
def segment_mean3(x: torch.Tensor, ptr: torch.Tensor) -> torch.Tensor:
    com = torch.empty(ptr.numel() - 1, 3, dtype=torch.float32, device=x.device)
    _lib.check(_L.pf_segment_mean3(_f(x), _i(ptr), ptr.numel() - 1, _f(com), _s()), "pf_segment_mean3")
    return com


@segment_mean3.register_fake
def _(x, ptr):
    return x.new_empty(ptr.numel() - 1, 3)


@torch.library.custom_op(f"{NS}::segment_shift3", mutates_args=("x",))
def segment_shift3(x: torch.Tensor, ptr: torch.Tensor, com: torch.Tensor, sign: float) -> None:
    _lib.check(_L.pf_segment_shift3(_f(x), _i(ptr), ptr.numel() - 1, _f(com), sign, _s()), "pf_segment_shift3")


# ------------------------------------------------------------------------------------------------ drivers
@torch.library.custom_op(f"{NS}::denoiser", mutates_args=("eps_h", "eps_x"))
def denoiser(eps_h: torch.Tensor, eps_x: torch.Tensor, args_addr: int) -> None:
    """One eps prediction over the buffers described by the PfSampleArgs at `args_addr`."""
    _lib.check(_L.pf_denoiser(C.c_void_p(args_addr), _s()), "pf_denoiser")


@torch.library.custom_op(f"{NS}::sample_loop", mutates_args=("pharm_x", "pharm_h", "prot_x"))
def sample_loop(pharm_x: torch.Tensor, pharm_h: torch.Tensor, prot_x: torch.Tensor, args_addr: int) -> None:
    """All reverse-diffusion steps described by the PfSampleArgs at `args_addr`, enqueued without returning to
    Python between steps."""
    _lib.check(_L.pf_sample_loop(C.c_void_p(args_addr), _s()), "pf_sample_loop")
