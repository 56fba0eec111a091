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
        # tcgen05 images of the message chains (only defined for the 3-GVP chain the kernel is built for)
        self.tc = None
        if n_msg == 3:
            imgs = [pack_message_tc(sd, f"{np_}.conv_layers.{l}", key) for l in range(n_convs) for key in ETYPE_KEYS]
            self.tc_stride = imgs[0].numel()
            assert self.tc_stride % 256 == 0
            self.tc = torch.cat(imgs).to(device)
        self.tcu = None
        if n_upd == 2:
            imgs = [pack_update_tc(sd, f"{np_}.conv_layers.{l}", nt) for l in range(n_convs) for nt in NTYPES]
            self.tcu_stride = imgs[0].numel()
            assert self.tcu_stride % 256 == 0
            self.tcu = torch.cat(imgs).to(device)

    def tc_ptr(self, conv: int, etype: int) -> int:
        return self.tc.data_ptr() + (conv * len(ETYPE_KEYS) + etype) * self.tc_stride

    def tcu_ptr(self, conv: int, ntype: int) -> int:
        return self.tcu.data_ptr() + (conv * len(NTYPES) + ntype) * self.tcu_stride

    def tcu_view(self, conv: int, ntype: int) -> torch.Tensor:
        o = (conv * len(NTYPES) + ntype) * self.tcu_stride
        return self.tcu[o:o + self.tcu_stride]

    def ptr(self, key: str) -> int:
        return self.flat.data_ptr() + 4 * self.offsets[key]

    def view(self, key: str) -> torch.Tensor:
        keys = list(self.offsets)
        i = keys.index(key)
        end = self.offsets[keys[i + 1]] if i + 1 < len(keys) else self.flat.numel()
        return self.flat[self.offsets[key]:end]


# ------------------------------------------------------------------------------------------------ tcgen05 operands
def split_bf16(w: torch.Tensor):
    """w ~= hi + lo with both parts bf16 (round to nearest even); 16 significand bits (tcgen05 self-test only)."""
    hi = w.float().to(torch.bfloat16)
    lo = (w.float() - hi.float()).to(torch.bfloat16)
    return hi, lo


def split_f16(w: torch.Tensor):
    """w ~= hi + lo with both parts fp16 (round to nearest even), as the kernels split activations on the fly:
    22 significand bits.  Weights beyond the fp16 range cannot be represented and are rejected."""
    if float(w.abs().max()) > 65000.0:
        raise ValueError("weight magnitude exceeds the fp16 range of the tcgen05 operand split")
    hi = w.float().to(torch.float16)
    lo = (w.float() - hi.float()).to(torch.float16)
    return hi, lo


def umma_b_image(w: torch.Tensor) -> torch.Tensor:
    """[N, K] bf16 weight (K contiguous = 'K-major') -> the SWIZZLE_NONE shared-memory image tcgen05.mma reads:
    8x8 core matrices of 128 contiguous bytes, element (n, k) at (k/8)*(N/8)*64 + (n/8)*64 + (n%8)*8 + k%8 (in
    bf16 units).  See csrc/pf_tc.cuh."""
    n, k = w.shape
    assert n % 8 == 0 and k % 8 == 0 and w.dtype in (torch.bfloat16, torch.float16)
    return w.reshape(n // 8, 8, k // 8, 8).permute(2, 0, 1, 3).contiguous().reshape(-1)


# --- K3 on the tensor cores: one byte image per (conv layer, edge type); layout constants mirror csrc/pf_tc_conv.cu
TC_SLAB_BYTES = 8192
TC_SLABS = (11, 9, 9)
TC_SMALL_OFF = sum(TC_SLABS) * TC_SLAB_BYTES
TC_GATE_OFF, TC_VEC_OFF, TC_CONST_OFF, TC_SMALL_BYTES = 0, 24576, 30720, 32768
TC_C_WH0, TC_C_WHU0, TC_C_WHC16 = 432, 452, 468
# The tcgen05 images carry Wf and bf multiplied by k = -log2(e): the kernels' SiLU then reads t = k (Wf s + bf)
# straight from the accumulator, SiLU(y) = t / (k (2^t + 1))  (csrc/pf_tc.cuh: silu_pre2).
TC_PRESCALE = -1.4426950408889634


def _bytes(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous().view(torch.uint8).reshape(-1)


def _hi_lo_images(w: torch.Tensor) -> torch.Tensor:
    """[N, 16] fp32 K-slab -> bytes of (hi image | lo image)."""
    hi, lo = split_f16(w)
    return torch.cat([_bytes(umma_b_image(hi)), _bytes(umma_b_image(lo))])


def pack_message_tc(sd: Dict[str, torch.Tensor], conv_p: str, etype_key: str) -> torch.Tensor:
    """The 3-GVP message chain of one edge type (gvp.py:392-415) as the uint8 image pf_edge_conv_tc streams:
    29 Wf^T K-slabs (fp16 hi | lo, UMMA SWIZZLE_NONE K-major) | gate images | [Wh | Wh.Wu] images | fp32 constants."""
    slabs, gates, vecs = [], [], []
    consts = torch.zeros((TC_SMALL_BYTES - TC_CONST_OFF) // 4, dtype=torch.float32)
    for g in range(3):
        q = f"{conv_p}.edge_message_fns.{etype_key}.{g}"
        Wh = sd[q + ".Wh"].detach().double().cpu()
        Wu = sd[q + ".Wu"].detach().double().cpu()
        Wf = sd[q + ".to_feats_out.0.weight"].detach().float().cpu()
        bf = sd[q + ".to_feats_out.0.bias"].detach().float().cpu()
        Wg = sd[q + ".scalar_to_vector_gates.weight"].detach().float().cpu()
        bg = sd[q + ".scalar_to_vector_gates.bias"].detach().float().cpu()
        vi, vh = Wh.shape
        assert Wf.shape == (128, 128 + (16 if g == 0 else 0) + vh) and Wg.shape == (16, 128) and Wu.shape == (vh, 16)
        assert (vi, vh) == ((17, 17) if g == 0 else (16, 16))
        K = 16 * TC_SLABS[g]
        Wp = torch.zeros(128, K)
        Wp[:, :Wf.shape[1]] = (Wf.double() * TC_PRESCALE).float()
        slabs += [_hi_lo_images(Wp[:, 16 * s:16 * s + 16]) for s in range(TC_SLABS[g])]
        gates += [_hi_lo_images(Wg[:, 16 * s:16 * s + 16]) for s in range(8)]
        Whu = Wh @ Wu                                        # [vi, 16], composed in float64
        k0 = 1 if g == 0 else 0                              # GVP 0: row 0 is the x_diff channel (CUDA cores)
        Bv = torch.cat([Wh[k0:k0 + 16, :16].t(), Whu[k0:k0 + 16, :].t()]).float()   # [32 (n), 16 (k)]
        vecs.append(_hi_lo_images(Bv))
        consts[144 * g:144 * g + 128] = (bf.double() * TC_PRESCALE).float()
        consts[144 * g + 128:144 * g + 144] = bg
        if g == 0:
            consts[TC_C_WH0:TC_C_WH0 + 17] = Wh[0, :].float()
            consts[TC_C_WHU0:TC_C_WHU0 + 16] = Whu[0, :].float()
            consts[TC_C_WHC16:TC_C_WHC16 + 16] = Wh[1:17, 16].float()
    blob = torch.cat(slabs + gates + vecs + [_bytes(consts)])
    assert blob.numel() == TC_SMALL_OFF + TC_SMALL_BYTES, blob.numel()
    return blob


# --- K4 on the tensor cores: one byte image per (conv layer, node type); mirrors Cfg<1> in csrc/pf_tc_conv.cu
TCU_SLABS = (9, 9)
TCU_SMALL_OFF = sum(TCU_SLABS) * TC_SLAB_BYTES
TCU_VEC_OFF, TCU_CONST_OFF, TCU_SMALL_BYTES = 16384, 20480, 24576
TCU_C_LN = 512          # ln_msg_w | ln_msg_b | ln_upd_w | ln_upd_b, 128 floats each


def pack_update_tc(sd: Dict[str, torch.Tensor], conv_p: str, ntype: str) -> torch.Tensor:
    """The 2-GVP node update of one node type (gvp.py:417-437, 511-532) as the uint8 image pf_node_update_tc
    streams: 18 Wf^T K-slabs | gate images | [Wh | Wh.Wu] images | fp32 constants (biases, LayerNorm rows)."""
    slabs, gates, vecs = [], [], []
    consts = torch.zeros((TCU_SMALL_BYTES - TCU_CONST_OFF) // 4, dtype=torch.float32)
    for g in range(2):
        q = f"{conv_p}.node_update_fns.{ntype}.{g}"
        Wh = sd[q + ".Wh"].detach().double().cpu()
        Wu = sd[q + ".Wu"].detach().double().cpu()
        Wf = sd[q + ".to_feats_out.0.weight"].detach().float().cpu()
        Wg = sd[q + ".scalar_to_vector_gates.weight"].detach().float().cpu()
        assert Wh.shape == (16, 16) and Wu.shape == (16, 16) and Wf.shape == (128, 144) and Wg.shape == (16, 128)
        Wfs = (Wf.double() * TC_PRESCALE).float()
        slabs += [_hi_lo_images(Wfs[:, 16 * s:16 * s + 16]) for s in range(9)]
        gates += [_hi_lo_images(Wg[:, 16 * s:16 * s + 16]) for s in range(8)]
        vecs.append(_hi_lo_images(torch.cat([Wh.t(), (Wh @ Wu).t()]).float()))
        consts[144 * g:144 * g + 128] = (sd[q + ".to_feats_out.0.bias"].detach().double().cpu() * TC_PRESCALE).float()
        consts[144 * g + 128:144 * g + 144] = sd[q + ".scalar_to_vector_gates.bias"].detach().float().cpu()
    for i, key in enumerate((f"{conv_p}.message_layer_norms.{ntype}.feat_norm.weight",
                             f"{conv_p}.message_layer_norms.{ntype}.feat_norm.bias",
                             f"{conv_p}.update_layer_norms.{ntype}.feat_norm.weight",
                             f"{conv_p}.update_layer_norms.{ntype}.feat_norm.bias")):
        consts[TCU_C_LN + 128 * i:TCU_C_LN + 128 * (i + 1)] = sd[key].detach().float().cpu()
    blob = torch.cat(slabs + gates + vecs + [_bytes(consts)])
    assert blob.numel() == TCU_SMALL_OFF + TCU_SMALL_BYTES, blob.numel()
    return blob
