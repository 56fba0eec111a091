"""ctypes binding of libpharmacoforge_b200.so (the C ABI declared in include/pharmacoforge_b200.h).

There is no CPU fallback: if the shared library is missing, import of any compute module fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PF_LIB_PATH: A/B builds of the same library (csrc/Makefile `variant`); experiments only
LIB_PATH = os.environ.get("PF_LIB_PATH") or os.path.join(_HERE, "libpharmacoforge_b200.so")

c_i32p = C.c_void_p
c_f32p = C.c_void_p
STREAM = C.c_void_p

# every symbol include/pharmacoforge_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pf_abi_version": (C.c_int, []),
    "pf_last_error": (C.c_char_p, []),
    "pf_gvp_layout": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "pf_scan_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "pf_exclusive_scan_i32": (C.c_int, [c_i32p, c_i32p, C.c_int64, C.c_void_p, C.c_size_t, STREAM]),
    "pf_radius_count": (C.c_int, [c_f32p, c_i32p, C.c_int32, C.c_float, C.c_int32, c_i32p, STREAM]),
    "pf_radius_fill": (C.c_int, [c_f32p, c_i32p, C.c_int32, C.c_float, C.c_int32, c_i32p, c_i32p, STREAM]),
    "pf_cell_radius_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "pf_cell_radius_count": (C.c_int, [c_f32p, c_i32p, C.c_int32, C.c_int64, C.c_float, C.c_int32, C.c_void_p, C.c_size_t,
                                       c_i32p, STREAM]),
    "pf_cell_radius_fill": (C.c_int, [c_f32p, c_i32p, C.c_int32, C.c_int64, C.c_float, C.c_int32, C.c_void_p, C.c_size_t,
                                      c_i32p, c_i32p, STREAM]),
    "pf_replicate_csr": (C.c_int, [c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_int32, c_i32p, c_i32p, c_i32p, STREAM]),
    "pf_dyn_graph": (C.c_int, [c_f32p, c_i32p, c_f32p, c_i32p, C.c_int32, C.c_float, C.c_int32, C.c_int32, c_i32p,
                               c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_void_p, STREAM]),
    "pf_dyn_graph_ffk": (C.c_int, [c_f32p, c_i32p, c_f32p, c_i32p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                   c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_void_p, STREAM]),
    "pf_dyn_graph_radius": (C.c_int, [c_f32p, c_i32p, c_f32p, c_i32p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_float,
                                      C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p,
                                      c_i32p, c_i32p, c_f32p, c_i32p, c_i32p, c_i32p, C.c_void_p, STREAM]),
    "pf_combine_subsegments": (C.c_int, [c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, C.c_int64, C.c_float, c_f32p, c_f32p, c_f32p,
                                         C.c_int32, STREAM]),
    "pf_plan_tiles": (C.c_int, [c_i32p, c_i32p, C.c_int32, C.c_int32, C.c_int32, c_i32p, C.c_int32, c_i32p, C.c_void_p,
                                STREAM]),
    "pf_plan_tiles3": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.c_int32, C.POINTER(C.c_void_p), C.c_int32, c_i32p, C.c_void_p, STREAM]),
    "pf_share_index": (C.c_int, [c_i32p, C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, STREAM]),
    "pf_share_gather": (C.c_int, [c_i32p, c_i32p, c_i32p, C.c_int32, C.c_int32, c_i32p, c_i32p, c_f32p, c_i32p, c_f32p, c_f32p,
                                  c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, STREAM]),
    "pf_plan_tiles_count": (C.c_int, [c_i32p, c_i32p, C.c_int32, C.c_int32, C.c_int32, c_i32p, C.c_void_p, STREAM]),
    "pf_plan_tiles_fill": (C.c_int, [c_i32p, c_i32p, C.c_int32, C.c_int32, C.c_int32, c_i32p, c_i32p, C.c_int32, c_i32p,
                                     C.c_void_p, STREAM]),
    "pf_tc_msg_blob_bytes": (C.c_size_t, []),
    "pf_tc_trace": (C.c_int, [C.c_void_p]),
    "pf_tc_upd_blob_bytes": (C.c_size_t, []),
    "pf_node_update_tc": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_void_p, c_f32p, c_f32p, STREAM]),
    "pf_edge_conv_tc": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p,
                                  C.c_int32, C.c_void_p, c_f32p, c_f32p, C.c_int32, STREAM]),
    "pf_node_update_tc_mapped": (C.c_int, [c_f32p, c_i32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_void_p, c_f32p, c_f32p,
                                           C.c_int32, STREAM]),
    "pf_edge_conv_tc_mapped": (C.c_int, [c_f32p, c_i32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p,
                                         c_i32p, C.c_int32, C.c_void_p, c_f32p, c_f32p, C.c_int32, C.c_int32, STREAM]),
    "pf_node_update_tc_f16": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_void_p, c_f32p, c_f32p, STREAM]),
    "pf_edge_conv_tc_f16": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p,
                                      C.c_int32, C.c_void_p, c_f32p, c_f32p, C.c_int32, STREAM]),
    "pf_seed_table": (C.c_int, [c_f32p, c_i32p, C.c_int32, c_f32p, c_f32p, STREAM]),
    "pf_edge_conv_tc_seeded": (C.c_int, [c_i32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p,
                                         C.c_int32, C.c_void_p, c_f32p, c_f32p, C.c_int32, C.c_int32, STREAM]),
    "pf_zero_i32": (C.c_int, [c_i32p, C.c_int64, STREAM]),
    "pf_fill_f32": (C.c_int, [c_f32p, C.c_int64, C.c_float, STREAM]),
    "pf_encode": (C.c_int, [c_f32p, C.c_int32, c_i32p, C.c_int32, c_f32p, c_f32p, c_f32p, STREAM]),
    "pf_edge_conv": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p,
                               C.c_int32, c_f32p, C.c_int32, c_f32p, c_f32p, C.c_int32, STREAM]),
    "pf_node_update": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, c_f32p, C.c_int32, c_f32p, c_f32p,
                                 STREAM]),
    "pf_noise_head": (C.c_int, [c_f32p, c_f32p, C.c_int64, c_f32p, C.c_int32, C.c_int32, c_f32p, c_f32p, STREAM]),
    "pf_posterior_step": (C.c_int, [c_f32p, c_f32p, C.c_int32, c_f32p, c_f32p, c_f32p, c_f32p, c_i32p, c_f32p,
                                    c_i32p, C.c_int32, C.c_float, C.c_float, C.c_float, STREAM]),
    "pf_posterior_step_ep": (C.c_int, [c_f32p, c_f32p, C.c_int32, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_uint32, c_i32p,
                                       c_f32p, c_i32p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_int32, STREAM]),
    "pf_posterior_step_philox": (C.c_int, [c_f32p, c_f32p, C.c_int32, c_f32p, c_f32p, C.c_void_p, C.c_uint32, c_i32p,
                                           c_f32p, c_i32p, C.c_int32, C.c_float, C.c_float, C.c_float, STREAM]),
    "pf_philox_normal": (C.c_int, [c_f32p, C.c_int64, C.c_void_p, C.c_uint32, C.c_uint32, STREAM]),
    "pf_segment_mean3": (C.c_int, [c_f32p, c_i32p, C.c_int32, c_f32p, STREAM]),
    "pf_segment_shift3": (C.c_int, [c_f32p, c_i32p, C.c_int32, c_f32p, C.c_float, STREAM]),
    "pf_sample_args_size": (C.c_size_t, []),
    "pf_tc_selftest": (C.c_int, [c_f32p, C.c_void_p, C.c_void_p, c_f32p, C.c_int32, C.c_int32, C.c_int32, STREAM]),
    "pf_launch_count": (C.c_int64, []),
    "pf_profile_enable": (C.c_int, [C.c_int32]),
    "pf_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int32]),
    "pf_train_sgemm": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                 C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, STREAM]),
    "pf_train_colsum": (C.c_int, [c_f32p, c_f32p, C.c_int64, C.c_int32, C.c_void_p, C.c_size_t, STREAM]),
    "pf_train_workspace_bytes": (C.c_size_t, []),
    "pf_train_gather_bwd_sorted": (C.c_int, [c_f32p, c_i32p, c_i32p, c_f32p, C.c_int64, C.c_int32, STREAM]),
    "pf_train_silu": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int64, STREAM]),
    "pf_train_gate": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int32, C.c_int32, STREAM]),
    "pf_train_vecnorm": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int32, STREAM]),
    "pf_train_layernorm_fwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int32, STREAM]),
    "pf_train_layernorm_bwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int32,
                                         C.c_void_p, C.c_size_t, STREAM]),
    "pf_train_vecln": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int32, STREAM]),
    "pf_train_gather": (C.c_int, [c_f32p, c_i32p, c_f32p, C.c_int64, C.c_int32, C.c_int32, STREAM]),
    "pf_train_segmean": (C.c_int, [c_f32p, c_i32p, c_i32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, STREAM]),
    "pf_train_edge_geom": (C.c_int, [c_f32p, c_f32p, c_i32p, c_i32p, c_f32p, c_f32p, C.c_int64, STREAM]),
    "pf_train_gvp_fwd": (C.c_int, [c_f32p] * 8 + [C.c_int64] + [C.c_int32] * 6 + [c_f32p] * 7 + [C.c_void_p, C.c_size_t, STREAM]),
    "pf_train_gvp_bwd": (C.c_int, [c_f32p] * 13 + [C.c_int64] + [C.c_int32] * 6 + [c_f32p] * 13 + [C.c_void_p, C.c_size_t, STREAM]),
    "pf_tc_gemm_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "pf_tc_gemm": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                             C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, STREAM]),
    "pf_train_gemm": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, STREAM]),
    "pf_scaled_accumulate": (C.c_int, [c_f32p, c_f32p, c_i32p, c_i32p, C.c_int64, C.c_float, c_f32p, c_f32p, c_f32p, C.c_int32,
                                       STREAM]),
    "pf_degree_norms": (C.c_int, [c_i32p, c_i32p, C.c_int32, C.c_int32, c_i32p, c_i32p, c_i32p, C.c_int32, c_f32p, c_f32p,
                                  STREAM]),
    "pf_denoiser": (C.c_int, [C.c_void_p, STREAM]),
    "pf_sample_loop": (C.c_int, [C.c_void_p, STREAM]),
}

MAX_CONVS = 8
ABI_VERSION = 5


class PfSampleArgs(C.Structure):
    """Mirror of `struct PfSampleArgs` in include/pharmacoforge_b200.h (field order matters)."""
    _fields_ = [
        ("n_graphs", C.c_int32), ("n_prot", C.c_int32), ("n_pharm", C.c_int32), ("n_prot_feats", C.c_int32),
        ("n_pharm_feats", C.c_int32),
        ("n_convs", C.c_int32), ("n_msg_gvps", C.c_int32), ("n_upd_gvps", C.c_int32), ("n_noise_gvps", C.c_int32),
        ("pf_k", C.c_int32), ("ff_max_nbrs", C.c_int32),
        ("ff_r", C.c_float),
        ("prot_x", C.c_void_p), ("prot_feats", C.c_void_p), ("prot_ptr", C.c_void_p),
        ("pharm_x", C.c_void_p), ("pharm_h", C.c_void_p), ("pharm_ptr", C.c_void_p),
        ("pp_start", C.c_void_p), ("pp_cnt", C.c_void_p), ("pp_col", C.c_void_p), ("pp_tiles", C.c_void_p),
        ("pp_n_tiles", C.c_void_p),
        ("pp_max_tiles", C.c_int32),
        ("ff_start", C.c_void_p),
        ("ff_cnt", C.c_void_p), ("ff_col", C.c_void_p), ("pf_start", C.c_void_p), ("pf_cnt", C.c_void_p),
        ("pf_col", C.c_void_p), ("fp_seg_dst", C.c_void_p), ("fp_seg_start", C.c_void_p),
        ("fp_seg_cnt", C.c_void_p), ("fp_col", C.c_void_p),
        ("pharm_chunk_ptr", C.c_void_p), ("fp_chunk_ptr", C.c_void_p),
        ("n_pharm_chunks", C.c_int32), ("n_fp_chunks", C.c_int32),
        ("ff_tiles", C.c_void_p), ("pf_tiles", C.c_void_p), ("fp_tiles", C.c_void_p), ("dyn_n_tiles", C.c_void_p),
        ("dyn_max_tiles", C.c_int32),
        ("prot_h", C.c_void_p), ("prot_v", C.c_void_p), ("prot_agg_h", C.c_void_p), ("prot_agg_v", C.c_void_p),
        ("pharm_hh", C.c_void_p), ("pharm_v", C.c_void_p), ("pharm_agg_h", C.c_void_p), ("pharm_agg_v", C.c_void_p),
        ("eps_h", C.c_void_p), ("eps_x", C.c_void_p),
        ("t_graph", C.c_void_p),
        ("w_pharm_enc", C.c_void_p), ("w_prot_enc", C.c_void_p),
        ("w_msg", (C.c_void_p * 4) * MAX_CONVS),
        ("w_upd", (C.c_void_p * 2) * MAX_CONVS),
        ("w_noise", C.c_void_p),
        ("w_msg_tc", (C.c_void_p * 4) * MAX_CONVS),
        ("w_upd_tc", (C.c_void_p * 2) * MAX_CONVS),
        ("tile_rows", C.c_int32),
        ("t_host", C.c_void_p), ("alpha_ts_host", C.c_void_p), ("var_terms_host", C.c_void_p),
        ("sigma_q_host", C.c_void_p),
        ("noise_x", C.c_void_p), ("noise_h", C.c_void_p),
        ("n_steps", C.c_int32),
        ("dev_status", C.c_void_p),
        ("flags", C.c_uint32),
        ("seed_row", C.c_void_p), ("seed_rep", C.c_void_p), ("seed_table", C.c_void_p), ("n_seed_rows", C.c_int32),
        ("noise_seed", C.c_void_p), ("noise_step0", C.c_int32), ("ff_k", C.c_int32),
        ("pk_x", C.c_void_p), ("pk_start", C.c_void_p), ("pk_cnt", C.c_void_p), ("pk_col", C.c_void_p),
        ("pk_tiles", C.c_void_p), ("pk_n_tiles", C.c_void_p), ("pk_max_tiles", C.c_int32), ("n_distinct", C.c_int32),
        ("pk_seed_row", C.c_void_p), ("pk_node0", C.c_void_p), ("enc_feats", C.c_void_p), ("enc_ptr", C.c_void_p),
        ("enc_rep", C.c_void_p), ("enc_table", C.c_void_p), ("aggd_h", C.c_void_p), ("aggd_v", C.c_void_p),
        ("c_x", C.c_void_p), ("c_h", C.c_void_p), ("c_v", C.c_void_p), ("c_agg_h", C.c_void_p), ("c_agg_v", C.c_void_p),
        ("c_seg_id", C.c_void_p), ("pf_col_c", C.c_void_p),
        ("ep_c1_host", C.c_void_p), ("ep_c2_host", C.c_void_p), ("ep_mode", C.c_int32),
        ("msg_norm_pharm", C.c_float), ("msg_norm_prot", C.c_float), ("tmp_agg_h", C.c_void_p), ("tmp_agg_v", C.c_void_p),
        ("pf_r", C.c_float), ("pf_max_nbrs", C.c_int32), ("pf_sub_ptr", C.c_void_p), ("fp_base", C.c_void_p),
        ("pf_sub_start", C.c_void_p), ("pf_sub_cnt", C.c_void_p), ("pf_sub_chunk_ptr", C.c_void_p),
        ("n_pf_sub_chunks", C.c_int32), ("n_pf_sub", C.c_int32), ("sub_agg_h", C.c_void_p), ("sub_agg_v", C.c_void_p),
        ("pf_sub_x", C.c_void_p),
        ("msg_norm_degree", C.c_int32), ("inv_norm_pharm", C.c_void_p), ("inv_norm_prot", C.c_void_p),
    ]


_lib = None


def load() -> C.CDLL:
    """dlopen the library and type every entry point; raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library has not been built (run `python -c 'import "
            f"__graft_entry__ as g; g.build()'` or `make -C pharmacoforge_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.pf_abi_version() != ABI_VERSION:
        raise ImportError("libpharmacoforge_b200.so ABI version mismatch")
    if lib.pf_sample_args_size() != C.sizeof(PfSampleArgs):
        raise ImportError("PfSampleArgs layout differs between _lib.py and the compiled library")
    _lib = lib
    return lib


class PfError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().pf_last_error().decode(errors="replace")
        raise PfError(f"{what} failed with code {rc}: {msg}")


DEV_STATUS_BITS = {
    1: "a destination node has more in-edges than one tile holds (64 rows FFMA / 128 rows tcgen05)",
    2: "a graph has more pharmacophore nodes than PF_MAX_PHARM_PER_GRAPH=128",
    4: "tile list capacity exceeded",
    8: "edge buffer capacity exceeded",
}


def check_dev_status(word: int):
    if word:
        msgs = [m for b, m in DEV_STATUS_BITS.items() if word & b]
        raise PfError("device status 0x%x: %s" % (word, "; ".join(msgs)))


PROFILE_SITES = ("dyn_graph", "plan", "encode", "edge_ff", "edge_pf", "edge_pp", "edge_fp", "update_pharm",
                 "update_prot", "noise_head", "posterior")


def profile_collect():
    """{site: (total_ms, launches)} since the last collect; call after synchronising the stream."""
    n = len(PROFILE_SITES)
    ms = (C.c_double * n)()
    cnt = (C.c_int32 * n)()
    check(load().pf_profile_collect(ms, cnt, n), "pf_profile_collect")
    return {s: (float(ms[i]), int(cnt[i])) for i, s in enumerate(PROFILE_SITES)}
