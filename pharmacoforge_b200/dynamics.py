"""Drop-in for the reference's `PharmRecDynamicsGVP` (pharmacoforge/models/dynamics_gvp.py:94-246).

Same constructor signature and the same `state_dict` keys and shapes, so checkpoints of the
reference load with `load_state_dict`.  The modules below only HOLD parameters under the reference's names; the
arithmetic runs in the CUDA library: `forward(g, timestep, batch_idxs)` packs the weights once (re-packed when
any parameter changes), binds the batch's device buffers into a `PfSampleArgs` and launches `pf_denoiser`.
"""
from __future__ import annotations

import os

import ctypes as C
import math
from typing import Dict, Optional, Union

import torch
import torch.nn as nn

from . import _lib, ops
from .batch import GraphBatch, on_batch_device
from .weights import PackedWeights

ALL_EDGES = [("pharm", "ff", "pharm"), ("prot", "pf", "pharm"), ("pharm", "fp", "prot"), ("prot", "pp", "prot")]


class GVP(nn.Module):
    """Parameter holder for one GVP; names and shapes as in the reference (gvp.py:43-87)."""

    def __init__(self, dim_vectors_in, dim_vectors_out, dim_feats_in, dim_feats_out, feats_activation=None,
                 vectors_activation=None):
        super().__init__()
        dim_h = max(dim_vectors_in, dim_vectors_out)
        self.Wh = nn.Parameter(torch.empty(dim_vectors_in, dim_h).uniform_(-1 / math.sqrt(dim_vectors_in),
                                                                          1 / math.sqrt(dim_vectors_in)))
        self.Wu = nn.Parameter(torch.empty(dim_h, dim_vectors_out).uniform_(-1 / math.sqrt(dim_h), 1 / math.sqrt(dim_h)))
        self.to_feats_out = nn.Sequential(nn.Linear(dim_h + dim_feats_in, dim_feats_out), nn.SiLU())
        self.scalar_to_vector_gates = nn.Linear(dim_feats_out, dim_vectors_out)
        self.vectors_activation = vectors_activation if vectors_activation is not None else nn.Sigmoid()


class _VDropout(nn.Module):
    def __init__(self, drop_rate):
        super().__init__()
        self.drop_rate = drop_rate
        self.dummy_param = nn.Parameter(torch.empty(0))


class GVPDropout(nn.Module):
    def __init__(self, rate):
        super().__init__()
        self.vector_dropout = _VDropout(rate)
        self.feat_dropout = nn.Dropout(rate)


class GVPLayerNorm(nn.Module):
    def __init__(self, feats_h_size, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.feat_norm = nn.LayerNorm(feats_h_size)


class GVPMultiEdgeConv(nn.Module):
    """Parameter holder for one heterogeneous conv layer (gvp.py:343-437)."""

    def __init__(self, etypes, scalar_size=128, vector_size=16, n_message_gvps=1, n_update_gvps=1, rbf_dim=16,
                 message_norm="mean", dropout=0.0):
        super().__init__()
        self.etypes = etypes
        self.message_norm = message_norm
        self.edge_message_fns = nn.ModuleDict()
        for et in etypes:
            gvps = []
            for i in range(n_message_gvps):
                vi = vector_size + 1 if i == 0 else vector_size
                si = scalar_size + rbf_dim if i == 0 else scalar_size
                gvps.append(GVP(vi, vector_size, si, scalar_size))
            self.edge_message_fns["_".join(et)] = nn.Sequential(*gvps)
        self.node_update_fns = nn.ModuleDict()
        self.update_layer_norms = nn.ModuleDict()
        self.message_layer_norms = nn.ModuleDict()
        for nt in sorted({et[2] for et in etypes}):
            self.node_update_fns[nt] = nn.Sequential(*[GVP(vector_size, vector_size, scalar_size, scalar_size)
                                                       for _ in range(n_update_gvps)])
            self.message_layer_norms[nt] = GVPLayerNorm(scalar_size)
            self.update_layer_norms[nt] = GVPLayerNorm(scalar_size)
        self.dropout = GVPDropout(dropout)


class NoisePredictionBlock(nn.Module):
    """dynamics_gvp.py:10-35."""

    def __init__(self, in_scalar_dim, out_scalar_dim, vector_size, n_gvps=3, intermediate_scalar_dim=64):
        super().__init__()
        gvps = []
        for i in range(n_gvps):
            last = i == n_gvps - 1
            gvps.append(GVP(vector_size, 1 if last else vector_size, in_scalar_dim,
                            intermediate_scalar_dim if last else in_scalar_dim,
                            vectors_activation=nn.Identity() if last else nn.Sigmoid()))
        self.gvps = nn.Sequential(*gvps)
        self.to_scalar_output = nn.Linear(intermediate_scalar_dim, out_scalar_dim)


class PharmRecGVP(nn.Module):
    """dynamics_gvp.py:44-82."""

    def __init__(self, in_scalar_dim, in_vector_dim, out_scalar_dim, n_convs=4, n_message_gvps=3, n_update_gvps=2,
                 message_norm="mean", n_noise_gvps=3, dropout=0.0):
        super().__init__()
        self.conv_layers = nn.ModuleList([
            GVPMultiEdgeConv(ALL_EDGES, in_scalar_dim, in_vector_dim, n_message_gvps, n_update_gvps,
                             message_norm=message_norm, dropout=dropout) for _ in range(n_convs)])
        self.noise_predictor = NoisePredictionBlock(in_scalar_dim, out_scalar_dim, in_vector_dim, n_noise_gvps)


class _DeviceState:
    """Feature / scratch buffers and the bound PfSampleArgs of one (module, batch) pair."""

    def __init__(self, dyn: "PharmRecDynamicsGVP", g: GraphBatch, w: PackedWeights):
        dev = g.device
        f32 = dict(dtype=torch.float32, device=dev)
        npn, nfn = max(g.n_prot, 1), max(g.n_pharm, 1)
        self.prot_h = torch.empty(npn, 128, **f32)
        self.prot_v = torch.empty(npn, 48, **f32)
        self.prot_agg_h = torch.empty(npn, 128, **f32)
        self.prot_agg_v = torch.empty(npn, 48, **f32)
        self.pharm_hh = torch.empty(nfn, 128, **f32)
        self.pharm_v = torch.empty(nfn, 48, **f32)
        self.pharm_agg_h = torch.empty(nfn, 128, **f32)
        self.pharm_agg_v = torch.empty(nfn, 48, **f32)
        self.eps_h = torch.empty(nfn, dyn.n_pharm_scalars, **f32)
        self.eps_x = torch.empty(nfn, 3, **f32)
        self.t_graph = torch.empty(max(g.n_graphs, 1), **f32)
        self.noise_seed = torch.zeros(1, dtype=torch.int64, device=dev)   # Philox key of the sampling loop (device-resident)
        self.graphs, self.warmed = {}, False                              # captured CUDA graphs of the loop
        if g.pharm_h is None or g.pharm_h.shape[1] != dyn.n_pharm_scalars:
            g.pharm_h = torch.zeros(nfn, dyn.n_pharm_scalars, **f32)
        a = _lib.PfSampleArgs()
        a.n_graphs, a.n_prot, a.n_pharm = g.n_graphs, g.n_prot, g.n_pharm
        a.n_prot_feats, a.n_pharm_feats = dyn.n_prot_scalars, dyn.n_pharm_scalars
        a.n_convs, a.n_msg_gvps, a.n_upd_gvps, a.n_noise_gvps = dyn.n_convs, dyn.n_message_gvps, dyn.n_update_gvps, dyn.n_noise_gvps
        if g.pf_k != dyn.pf_k:
            raise ValueError(f"the batch was built for pf_k = {g.pf_k}, the model has pf_k = {dyn.pf_k} "
                             "(GraphBatch.from_pockets(pf_k=...))")
        a.pf_k, a.ff_max_nbrs, a.ff_r = g.pf_k, g.ff_max_nbrs, float(dyn.graph_cutoffs["ff"])
        a.ff_k = int(dyn.ff_k)
        for name in ("prot_x", "prot_feats", "prot_ptr", "pharm_x", "pharm_h", "pharm_ptr", "pp_start", "pp_cnt",
                     "pp_col", "pp_tiles", "pp_n_tiles", "ff_start", "ff_cnt", "ff_col", "pf_start", "pf_cnt",
                     "pf_col", "fp_seg_dst", "fp_seg_start", "fp_seg_cnt", "fp_col", "pharm_chunk_ptr",
                     "fp_chunk_ptr", "ff_tiles", "pf_tiles", "fp_tiles", "dyn_n_tiles"):
            setattr(a, name, getattr(g, name).data_ptr())
        for name in ("prot_h", "prot_v", "prot_agg_h", "prot_agg_v", "pharm_hh", "pharm_v", "pharm_agg_h",
                     "pharm_agg_v", "eps_h", "eps_x", "t_graph"):
            setattr(a, name, getattr(self, name).data_ptr())
        a.pp_max_tiles = g.pp_num_tiles
        a.n_pharm_chunks, a.n_fp_chunks = g.n_chunks, g.n_fp_chunks
        if g.pf_k == 0:       # radius pf / fp edges: sub-segment list of the pf segments + scratch for one mean per sub-segment
            a.pf_r, a.pf_max_nbrs = float(dyn.graph_cutoffs["pf"]), g.pf_max_nbrs
            for name in ("pf_sub_ptr", "fp_base", "pf_sub_start", "pf_sub_cnt", "pf_sub_chunk_ptr", "pf_sub_x"):
                setattr(a, name, getattr(g, name).data_ptr())
            a.n_pf_sub_chunks, a.n_pf_sub = g.n_graphs, g.n_pf_sub
            self.sub_agg_h = torch.empty(max(g.n_pf_sub, 1), 128, **f32)
            self.sub_agg_v = torch.empty(max(g.n_pf_sub, 1), 48, **f32)
            a.sub_agg_h, a.sub_agg_v = self.sub_agg_h.data_ptr(), self.sub_agg_v.data_ptr()
        a.dyn_max_tiles = g.dyn_max_tiles
        a.dev_status = g.status.data_ptr()
        a.w_pharm_enc, a.w_prot_enc, a.w_noise = w.ptr("pharm_enc"), w.ptr("prot_enc"), w.ptr("noise")
        for l in range(dyn.n_convs):
            for e in range(4):
                a.w_msg[l][e] = w.ptr(f"msg{l}_{e}")
            for n in range(2):
                a.w_upd[l][n] = w.ptr(f"upd{l}_{n}")
        a.tile_rows = g.tile_rows
        # first-layer seeding (one-hot protein features only): static row / representative tables, per-call scratch
        self.seed = None
        if g.tile_rows == 128 and g.n_prot > 0 and g.n_pp_edges > 0:
            self.seed = g.seed_arrays()
        if self.seed is not None:
            seed_row, seed_rep = self.seed
            self.seed_table = torch.zeros(seed_rep.numel(), 128, **f32)
            a.seed_row, a.seed_rep, a.seed_table = seed_row.data_ptr(), seed_rep.data_ptr(), self.seed_table.data_ptr()
            a.n_seed_rows = seed_rep.numel()
            # first-layer encoder table (PF_FLAG_NO_LAYER0_TABLE in the header): the protein scalars of the first conv layer
            # are read from one row per (graph, atom type) instead of a materialised [n_prot][128] array
            self.enc = g.enc_arrays()
            self.enc_table = torch.zeros(seed_rep.numel(), 128, **f32)
            a.enc_feats, a.enc_ptr, a.enc_rep = (self.enc[k].data_ptr() for k in ("enc_feats", "enc_ptr", "enc_rep"))
            a.enc_table = self.enc_table.data_ptr()
        if dyn.message_norm != "mean":      # numeric message_norm: scratch for one edge type's means (see the header)
            rows = max(npn, nfn)
            self.tmp_agg_h = torch.empty(rows, 128, **f32)
            self.tmp_agg_v = torch.empty(rows, 48, **f32)
            a.tmp_agg_h, a.tmp_agg_v = self.tmp_agg_h.data_ptr(), self.tmp_agg_v.data_ptr()
            a.msg_norm_pharm = a.msg_norm_prot = float(dyn.message_norm)
            if dyn.message_norm == 0:       # per-graph divisor (edges per node + 1), one reciprocal per node
                if g.pf_k > 0 and g.n_pharm > g.n_prot:
                    raise ValueError("message_norm = 0 with kNN pf edges needs n_pharm <= n_prot: the reference looks the "
                                     "pharmacophore node index up in the protein batch index (dynamics_gvp.py:220)")
                self.inv_norm_pharm = torch.empty(nfn, **f32)
                self.inv_norm_prot = torch.empty(npn, **f32)
                a.msg_norm_degree = 1
                a.inv_norm_pharm, a.inv_norm_prot = self.inv_norm_pharm.data_ptr(), self.inv_norm_prot.data_ptr()
        self.share = None     # buffers of the shared-pocket mode, bound on first use (bind_share)
        if g.tile_rows == 128:
            if w.tc is None:
                raise NotImplementedError("the tcgen05 message kernel is built for n_message_gvps=3 (configs/dev.yml); "
                                          "build the batch with tile_rows=64 to use the fp32 FFMA kernels")
            for l in range(dyn.n_convs):
                for e in range(4):
                    a.w_msg_tc[l][e] = w.tc_ptr(l, e)
                if w.tcu is not None:
                    for n in range(2):
                        a.w_upd_tc[l][n] = w.tcu_ptr(l, n)
        self.args = a
        self.weights = w          # keep alive
        self.batch_buffers = (g.prot_x, g.pharm_x, g.pharm_h)

    def bind_share(self, g: GraphBatch) -> bool:
        """Bind the static arrays and scratch of the shared-pocket mode; False when the batch does not qualify."""
        if self.share is not None:
            return True
        sh = g.share_arrays()
        if sh is None or self.seed is None:
            return False
        dev, a = g.device, self.args
        f32 = dict(dtype=torch.float32, device=dev)
        n_c, n_d, rows = max(g.pf_k * g.n_pharm, 1), sh["n_distinct"], sh["enc_rep"].numel()
        buf = dict(enc_table=self.enc_table, aggd_h=torch.zeros(n_d, 128, **f32),
                   aggd_v=torch.zeros(n_d, 48, **f32), c_x=torch.zeros(n_c, 3, **f32), c_h=torch.zeros(n_c, 128, **f32),
                   c_v=torch.zeros(n_c, 48, **f32), c_agg_h=torch.zeros(n_c, 128, **f32), c_agg_v=torch.zeros(n_c, 48, **f32),
                   c_seg_id=torch.arange(n_c, dtype=torch.int32, device=dev),
                   pf_col_c=torch.zeros(n_c, dtype=torch.int32, device=dev))
        for name in ("pk_x", "pk_start", "pk_cnt", "pk_col", "pk_tiles", "pk_n_tiles", "pk_seed_row", "pk_node0", "enc_feats",
                     "enc_ptr", "enc_rep"):
            setattr(a, name, sh[name].data_ptr())
        for name, t in buf.items():
            setattr(a, name, t.data_ptr())
        a.pk_max_tiles, a.n_distinct = sh["pk_max_tiles"], n_d
        if self.seed_table.shape[0] < rows:      # the table is indexed by (graph, type) in both modes: same size
            raise AssertionError("seed table smaller than the encoder table")
        self.share = (sh, buf)                   # keep alive
        return True

    @property
    def addr(self) -> int:
        return C.addressof(self.args)


class PharmRecDynamicsGVP(nn.Module):
    """`self.dynamics` of the diffusion model.  forward(g, timestep, batch_idxs) -> (eps_h [Nf,6], eps_x [Nf,3])."""

    def __init__(self, n_pharm_scalars, n_prot_scalars, vector_size: int = 16, n_convs=4, n_hidden_scalars=128,
                 act_fn=nn.SiLU, message_norm: Union[float, str, Dict] = 1, graph_cutoffs: dict = {},
                 n_message_gvps: int = 3, n_update_gvps: int = 2, n_noise_gvps: int = 3, dropout: float = 0.0,
                 ff_k: int = 0, pf_k: int = 0):
        super().__init__()
        if vector_size != 16 or n_hidden_scalars != 128:
            raise NotImplementedError("the sm_100a kernels are built for vector_size=16, n_hidden_scalars=128 "
                                      "(configs/dev.yml); other widths need a rebuild with new tile constants")
        # message_norm (gvp.py:375-389): 'mean' (configs/dev.yml) = mean over the in-edges per edge type; a positive number =
        # SUM over the in-edges divided by it (the reference constructor's default is 1); 0 = SUM divided by the graph's edges
        # per node + 1 (:504-507, pf_degree_norms).  A dict raises upstream too (check_message_norm calls .keys() on a set, :453).
        if isinstance(message_norm, str):
            if message_norm != "mean":
                raise ValueError(f"invalid message_norm {message_norm!r}")
        elif isinstance(message_norm, (int, float)) and not isinstance(message_norm, bool):
            if message_norm < 0:
                raise ValueError("message_norm must be >= 0")
        else:
            raise NotImplementedError("message_norm must be 'mean' or a number >= 0 (a dict fails in the reference's own "
                                      "check_message_norm, gvp.py:453)")
        self.message_norm = message_norm
        # pf_k == 0 (the reference constructor's default; configs/dev.yml uses 5): pf / fp edges from radius(pharm, prot,
        # r = graph_cutoffs['pf'], 100 per protein atom) instead of kNN (dynamics_gvp.py:210-216) -- pf_dyn_graph_radius
        if pf_k < 0:
            raise ValueError("pf_k must be >= 0")
        if pf_k == 0 and "pf" not in graph_cutoffs:
            raise KeyError("pf_k == 0 reads graph_cutoffs['pf'] (dynamics_gvp.py:211)")
        if ff_k < 0:
            raise ValueError("ff_k must be >= 0")
        if act_fn is not nn.SiLU:
            raise NotImplementedError("only SiLU is built")
        self.graph_cutoffs = graph_cutoffs
        self.n_pharm_scalars, self.n_prot_scalars = n_pharm_scalars, n_prot_scalars
        self.vector_size, self.ff_k, self.pf_k = vector_size, ff_k, pf_k
        self.n_convs, self.n_message_gvps, self.n_update_gvps, self.n_noise_gvps = n_convs, n_message_gvps, n_update_gvps, n_noise_gvps
        self.pharm_encoder = nn.Sequential(nn.Linear(n_pharm_scalars + 1, n_hidden_scalars), act_fn(),
                                           nn.LayerNorm(n_hidden_scalars))
        self.prot_encoder = nn.Sequential(nn.Linear(n_prot_scalars + 1, n_hidden_scalars), act_fn(),
                                          nn.LayerNorm(n_hidden_scalars))
        self.noise_predictor = PharmRecGVP(n_hidden_scalars, vector_size, n_pharm_scalars, n_convs, n_message_gvps,
                                           n_update_gvps, message_norm, n_noise_gvps, dropout)
        # Opt-in exact dead-work elimination: the protein-side outputs of the last conv layer are never read
        # (dynamics_gvp.py:84-92 feeds only node_data['pharm'] to the noise head), so their pp / fp messages and
        # protein node update can be dropped without changing a bit of (eps_h, eps_x).  Off by default: the nominal
        # path does the reference's full work.  Not a constructor argument (the reference signature is kept).
        self.skip_dead_work = False
        # Edge / update MLP precision on the tensor cores: "fp32" (default; fp16 hi/lo split, three passes, inside the
        # 1e-4 parity bar) or "fp16" (PF_FLAG_FP16_SINGLE_PASS: one pass over 11-bit operands, SiLU on packed fp16
        # pairs -- the reduced-precision path of BASELINE.json configs[3], tolerance stated in the tests).
        self.edge_mlp_precision = "fp32"
        # First conv layer, pp edges: the per-node part of GVP 0's scalar contraction (Wf0[:, :128] h_src) comes from a
        # (graph, atom type) table instead of a per-edge gather + contraction (SURVEY.md hard part 2's exact split; needs
        # one-hot protein features, checked per batch).  False = the general kernel, the A/B switch of the parity tests.
        self.layer0_seed = True
        # First conv layer, protein scalars: one encoder row per (graph, atom type) read through a row map by the pf
        # message gather and the protein node update, instead of a per-node encoder pass and a [n_prot][128] array
        # (PF_FLAG_NO_LAYER0_TABLE; same values, bit-identical results; one-hot protein features, with layer0_seed).
        self.layer0_table = os.environ.get("PF_LAYER0_TABLE", "1") != "0"   # env: A/B switch for measurements
        # Opt-in exact work elimination for SAMPLING (SURVEY.md hard part 5a + 5b + 5c, csrc/pf_share.cu): first-layer pp
        # messages once per distinct pocket, protein rows encoded / updated only where the last layer reads them.  Needs
        # the same timestep for every graph (the reverse-diffusion loop; forward() checks it) and one-hot protein
        # features; equals the nominal path up to fp32 rounding of x_src - x_dst.  Never the default, never the headline.
        self.share_pocket_messages = False
        # Direct callers of forward() get the device status word checked on every call (one 4-byte D2H sync);
        # PharmacophoreDiff's own loops check once at their end and switch this off around their calls.
        self.check_status_every_call = True
        self._packed: Optional[PackedWeights] = None
        self._packed_key = None

    # ------------------------------------------------------------------ weights
    def packed_weights(self, device) -> PackedWeights:
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._packed_key != key:
            sd = {"dyn." + k: v for k, v in self.state_dict().items()}
            self._packed = PackedWeights(sd, "dyn", self.n_convs, self.n_message_gvps, self.n_update_gvps,
                                         self.n_noise_gvps, device)
            self._packed_key = key
        return self._packed

    def bind(self, g: GraphBatch) -> _DeviceState:
        """Device buffers + argument block for batch `g` (cached on the batch)."""
        w = self.packed_weights(g.device)
        st = getattr(g, "_pf_state", None)
        cur = (g.prot_x, g.pharm_x, g.pharm_h)
        if st is None or st.weights is not w or any(a is not b for a, b in zip(st.batch_buffers, cur)):
            st = _DeviceState(self, g, w)
            g._pf_state = st
        if self.edge_mlp_precision not in ("fp32", "fp16"):
            raise ValueError("edge_mlp_precision must be 'fp32' or 'fp16'")
        share = False
        if self.share_pocket_messages:
            if self.message_norm != "mean":
                raise NotImplementedError("share_pocket_messages is built for message_norm = 'mean'")
            if self.pf_k == 0:
                raise NotImplementedError("share_pocket_messages is built for pf_k >= 1 (kNN pf edges)")
            if self.n_convs != 2 or self.n_message_gvps != 3 or self.n_update_gvps != 2 or g.tile_rows != 128:
                raise NotImplementedError("share_pocket_messages is built for the dev.yml shape (n_convs=2, 3 message / 2 "
                                          "update GVPs) on the tcgen05 path")
            share = st.bind_share(g)             # False: the batch does not qualify (features not one-hot): nominal path
        st.args.flags = ((1 if (self.skip_dead_work or share) else 0) |     # PF_FLAG_SKIP_DEAD_WORK
                         (2 if self.edge_mlp_precision == "fp16" else 0) |  # PF_FLAG_FP16_SINGLE_PASS
                         (0 if self.layer0_seed else 4) |                   # PF_FLAG_NO_LAYER0_SEED
                         (8 if share else 0) |                              # PF_FLAG_SHARE_POCKET_MESSAGES
                         (0 if self.layer0_table else 16))                  # PF_FLAG_NO_LAYER0_TABLE
        return st

    # ------------------------------------------------------------------ forward
    @on_batch_device
    def forward(self, g: GraphBatch, timestep: torch.Tensor, batch_idxs=None):
        if self.training and self.noise_predictor.conv_layers[0].dropout.feat_dropout.p > 0:
            raise NotImplementedError("the fused kernels implement eval-mode semantics (no dropout, no autograd graph): call "
                                      ".eval(), or train through PharmacophoreDiff.training_step (train_graph.py)")
        st = self.bind(g)
        tt = timestep.to(device=g.device, dtype=torch.float32).reshape(-1)
        if (st.args.flags & 8) and tt.numel() > 1 and not bool((tt == tt[0]).all().item()):
            raise ValueError("share_pocket_messages needs the same timestep for every graph of the batch")
        st.t_graph.copy_(tt)
        ops.denoiser(st.eps_h, st.eps_x, st.addr)
        if self.check_status_every_call:
            g.check_status()   # device-detected conditions (degree / tile / edge overflow) must not pass silently
        return st.eps_h, st.eps_x
