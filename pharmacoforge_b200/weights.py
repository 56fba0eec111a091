"""Packing of reference-layout state_dict tensors into the flat fp32 blocks the kernels read.

The layout of one packed GVP is defined once, in C (pf_gvp_layout, include/pharmacoforge_b200.h); this
module queries it instead of restating it.  Keys follow SURVEY.md App. C / the reference's state_dict.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib

ETYPE_KEYS = ("pharm_ff_pharm", "prot_pf_pharm", "pharm_fp_prot", "prot_pp_prot")  # ff, pf, fp, pp
NTYPES = ("pharm", "prot")


def gvp_layout(vi: int, vo: int, si: int, so: int):
    offs = (C.c_int64 * 6)()
    total = _lib.load().pf_gvp_layout(vi, vo, si, so, offs)
    return list(offs), int(total)


def pack_gvp(sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """One GVP (reference gvp.py:43-116 parameters Wh, Wu, to_feats_out.0, scalar_to_vector_gates)."""
    Wh = sd[p + ".Wh"].detach().float().cpu()
    Wu = sd[p + ".Wu"].detach().float().cpu()
    Wf = sd[p + ".to_feats_out.0.weight"].detach().float().cpu()
    bf = sd[p + ".to_feats_out.0.bias"].detach().float().cpu()
    Wg = sd[p + ".scalar_to_vector_gates.weight"].detach().float().cpu()
    bg = sd[p + ".scalar_to_vector_gates.bias"].detach().float().cpu()
    vi, vh = Wh.shape
    vo = Wu.shape[1]
    so, k = Wf.shape
    si = k - vh
    assert Wu.shape[0] == vh and vh == max(vi, vo) and Wg.shape == (vo, so)
    offs, total = gvp_layout(vi, vo, si, so)
    out = torch.zeros(total, dtype=torch.float32)
    out[offs[0]:offs[0] + vi * vh] = Wh.reshape(-1)
    out[offs[1]:offs[1] + vh * vo] = Wu.reshape(-1)
    out[offs[2]:offs[2] + k * so] = Wf.t().contiguous().reshape(-1)   # [K][so]; padding rows stay zero
    out[offs[3]:offs[3] + so] = bf
    out[offs[4]:offs[4] + so * vo] = Wg.t().contiguous().reshape(-1)  # [so][vo]
    out[offs[5]:offs[5] + vo] = bg
    return out


def pack_encoder(sd, p: str) -> torch.Tensor:
    """Sequential(Linear, SiLU, LayerNorm) of dynamics_gvp.py:107-117 -> Wt[(nf+1)][128] | b | ln_w | ln_b."""
    W = sd[p + ".0.weight"].detach().float().cpu()
    return torch.cat([W.t().contiguous().reshape(-1), sd[p + ".0.bias"].detach().float().cpu(),
                      sd[p + ".2.weight"].detach().float().cpu(), sd[p + ".2.bias"].detach().float().cpu()])


def pack_message(sd, conv_p: str, etype_key: str, n_gvps: int) -> torch.Tensor:
    return torch.cat([pack_gvp(sd, f"{conv_p}.edge_message_fns.{etype_key}.{i}") for i in range(n_gvps)])


def pack_update(sd, conv_p: str, ntype: str, n_gvps: int) -> torch.Tensor:
    parts = [sd[f"{conv_p}.message_layer_norms.{ntype}.feat_norm.weight"], sd[f"{conv_p}.message_layer_norms.{ntype}.feat_norm.bias"],
             sd[f"{conv_p}.update_layer_norms.{ntype}.feat_norm.weight"], sd[f"{conv_p}.update_layer_norms.{ntype}.feat_norm.bias"]]
    parts = [t.detach().float().cpu() for t in parts]
    parts += [pack_gvp(sd, f"{conv_p}.node_update_fns.{ntype}.{i}") for i in range(n_gvps)]
    return torch.cat(parts)


def pack_noise_head(sd, p: str, n_gvps: int) -> torch.Tensor:
    parts = [pack_gvp(sd, f"{p}.gvps.{i}") for i in range(n_gvps)]
    W = sd[p + ".to_scalar_output.weight"].detach().float().cpu()   # [n_out, 64]
    b = sd[p + ".to_scalar_output.bias"].detach().float().cpu()
    n_out, k = W.shape
    n4 = (n_out + 3) // 4 * 4
    Wt = torch.zeros(k, n4)
    Wt[:, :n_out] = W.t()
    bb = torch.zeros(n4)
    bb[:n_out] = b
    return torch.cat(parts + [Wt.reshape(-1), bb])


class PackedWeights:
    """All kernel weight blocks of one PharmRecDynamicsGVP, resident on one device."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, n_convs: int, n_msg: int, n_upd: int, n_noise: int,
                 device):
        np_ = f"{prefix}.noise_predictor"
        blocks = {"pharm_enc": pack_encoder(sd, f"{prefix}.pharm_encoder"),
                  "prot_enc": pack_encoder(sd, f"{prefix}.prot_encoder")}
        for l in range(n_convs):
            cp = f"{np_}.conv_layers.{l}"
            for e, key in enumerate(ETYPE_KEYS):
                blocks[f"msg{l}_{e}"] = pack_message(sd, cp, key, n_msg)
            for n, nt in enumerate(NTYPES):
                blocks[f"upd{l}_{n}"] = pack_update(sd, cp, nt, n_upd)
        blocks["noise"] = pack_noise_head(sd, f"{np_}.noise_predictor", n_noise)
        # one allocation; every block starts on a 16-byte boundary (all sizes are multiples of 4 floats)
        offs, cur = {}, 0
        for k, t in blocks.items():
            assert t.numel() % 4 == 0, k
            offs[k] = cur
            cur += t.numel()
        self.flat = torch.cat(list(blocks.values())).to(device)
        self.offsets = offs
        self.n_convs = n_convs

    def ptr(self, key: str) -> int:
        return self.flat.data_ptr() + 4 * self.offsets[key]

    def view(self, key: str) -> torch.Tensor:
        keys = list(self.offsets)
        i = keys.index(key)
        end = self.offsets[keys[i + 1]] if i + 1 < len(keys) else self.flat.numel()
        return self.flat[self.offsets[key]:end]


# ------------------------------------------------------------------------------------------------ tcgen05 operands
def split_bf16(w: torch.Tensor):
    """w ~= hi + lo with both parts bf16 (round to nearest even), as the kernels split activations on the fly."""
    hi = w.float().to(torch.bfloat16)
    lo = (w.float() - hi.float()).to(torch.bfloat16)
    return hi, lo


def umma_b_image(w: torch.Tensor) -> torch.Tensor:
    """[N, K] bf16 weight (K contiguous = 'K-major') -> the SWIZZLE_NONE shared-memory image tcgen05.mma reads:
    8x8 core matrices of 128 contiguous bytes, element (n, k) at (k/8)*(N/8)*64 + (n/8)*64 + (n%8)*8 + k%8 (in
    bf16 units).  See csrc/pf_tc.cuh."""
    n, k = w.shape
    assert n % 8 == 0 and k % 8 == 0 and w.dtype == torch.bfloat16
    return w.reshape(n // 8, 8, k // 8, 8).permute(2, 0, 1, 3).contiguous().reshape(-1)
