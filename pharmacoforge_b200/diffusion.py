"""Drop-in for the reference's `PharmacophoreDiff` (pharmacoforge/models/pharmacodiff.py:25-578): same constructor
arguments (fed by `model_from_config`, config_utils/load_from_config.py:16-30), same `state_dict` layout
(`gamma.gamma`, `dynamics.*`), same sampling entry points; the denoiser and the reverse-diffusion loop run in
the CUDA library.  No pytorch_lightning dependency: `load_from_checkpoint` reads the Lightning checkpoint dict
(`hyper_parameters`, `state_dict`) directly.
"""
from __future__ import annotations

import ctypes as C
from math import ceil
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .batch import GraphBatch, Pocket, on_batch_device
from .dynamics import PharmRecDynamicsGVP
from .hostutil import polynomial_gamma  # noqa: F401  (re-exported: the schedule is host-only code)


class PredefinedNoiseSchedule(nn.Module):
    """Lookup table gamma[round(t*T)] (pharmacodiff.py:636-668); `gamma` is a frozen Parameter so that it sits
    in the state_dict under `gamma.gamma` like the reference's."""

    def __init__(self, noise_schedule: str, timesteps: int, precision: float):
        super().__init__()
        self.timesteps = timesteps
        kind, _, power = noise_schedule.partition("_")
        if kind != "polynomial" or not power:
            raise ValueError(noise_schedule)
        self.gamma = nn.Parameter(polynomial_gamma(timesteps, precision, float(power)), requires_grad=False)

    def forward(self, t):
        return self.gamma[torch.round(t * self.timesteps).long()]


class SampledPharmacophore:
    """Result object (analysis/pharm_builder.py:7-71): coordinates, argmax types, xyz writer."""

    type_idx_to_elem = ["P", "S", "F", "N", "O", "C"]

    def __init__(self, ph_coords: torch.Tensor, ph_feats: torch.Tensor, pharm_type_map: List[str], traj_frames=None):
        self.pharm_type_map = pharm_type_map
        self.ph_coords = ph_coords
        self.ph_feats = ph_feats
        self.ph_feats_idxs = ph_feats.argmax(dim=1)
        self.ph_types = [pharm_type_map[int(i)] for i in self.ph_feats_idxs]
        self.n_ph_centers = ph_coords.shape[0]
        self.pos_frames, self.feat_frames = traj_frames if traj_frames is not None else (None, None)
        self.ph_type_to_elem = {pharm_type_map[i]: self.type_idx_to_elem[i] for i in range(len(pharm_type_map))}

    def pharm_to_xyz(self, pos, types):
        lines = [f"{len(pos)}"]
        for i in range(len(pos)):
            lines.append(f"{self.ph_type_to_elem[types[i]]} {pos[i, 0]:.3f} {pos[i, 1]:.3f} {pos[i, 2]:.3f}")
        return "\n".join(lines) + "\n"

    def to_xyz_file(self, filename: Optional[str] = None):
        out = self.pharm_to_xyz(self.ph_coords, self.ph_types)
        if filename is None:
            return out
        Path(filename).write_text(out)

    def traj_to_xyz(self, filename: Optional[str] = None):
        if self.pos_frames is None:
            raise ValueError("no trajectory frames were recorded for this pharmacophore")
        idx = self.feat_frames.argmax(dim=2)
        out = "".join(self.pharm_to_xyz(self.pos_frames[i], [self.pharm_type_map[int(j)] for j in idx[i]])
                      for i in range(self.pos_frames.shape[0]))
        if filename is None:
            return out
        Path(filename).write_text(out)


class PharmacophoreDiff(nn.Module):
    def __init__(self, pharm_nf, rec_nf, ph_type_map: List[str], processed_data_dir=None, n_timesteps: int = 1000,
                 graph_config={}, dynamics_config={}, lr_scheduler_config={}, sample_interval: float = 1,
                 val_loss_interval: float = 1, batch_size: int = 64, pharms_per_pocket: int = 8,
                 n_pockets_to_sample: int = 8, precision=1e-4, pharm_feat_norm_constant=1,
                 endpoint_param_feat: bool = False, endpoint_param_coord: bool = False, weighted_loss: bool = False,
                 remove_com: bool = True, **kwargs):
        super().__init__()
        self.hparams = dict(pharm_nf=pharm_nf, rec_nf=rec_nf, ph_type_map=ph_type_map,
                            processed_data_dir=processed_data_dir, n_timesteps=n_timesteps, graph_config=graph_config,
                            dynamics_config=dynamics_config, lr_scheduler_config=lr_scheduler_config,
                            sample_interval=sample_interval, val_loss_interval=val_loss_interval,
                            batch_size=batch_size, pharms_per_pocket=pharms_per_pocket,
                            n_pockets_to_sample=n_pockets_to_sample, precision=precision,
                            pharm_feat_norm_constant=pharm_feat_norm_constant,
                            endpoint_param_feat=endpoint_param_feat, endpoint_param_coord=endpoint_param_coord,
                            weighted_loss=weighted_loss, remove_com=remove_com, **kwargs)
        self.n_pharm_feats, self.n_prot_feats = pharm_nf, rec_nf
        self.batch_size = batch_size
        self.ph_type_map = ph_type_map
        self.n_timesteps = n_timesteps
        self.remove_com = remove_com
        self.endpoint_param_feat, self.endpoint_param_coord = bool(endpoint_param_feat), bool(endpoint_param_coord)
        self.pharm_feat_norm_constant = pharm_feat_norm_constant
        self.weighted_loss = weighted_loss
        self.gamma = PredefinedNoiseSchedule("polynomial_2", n_timesteps, precision)
        self.dynamics = PharmRecDynamicsGVP(pharm_nf, rec_nf, **graph_config, **dynamics_config)
        self.lr_scheduler_config = lr_scheduler_config
        self.sample_interval, self.val_loss_interval = sample_interval, val_loss_interval
        self.pharms_per_pocket, self.n_pockets_to_sample = pharms_per_pocket, n_pockets_to_sample
        self.graph_cutoffs = graph_config.get("graph_cutoffs", {})
        self._tables = None
        # Replay the T-step loop as one captured CUDA graph when the noise comes from the in-kernel Philox generator
        # (sample_given_receptor(noise=None)); not a constructor argument (the reference signature is kept).
        self.use_cuda_graph = False

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_config(cls, config: dict) -> "PharmacophoreDiff":
        """config_utils/load_from_config.py:6-32 (model_from_config)."""
        ev = config["training"]["evaluation"]
        return cls(pharm_nf=len(config["dataset"]["ph_type_map"]), rec_nf=len(config["dataset"]["prot_elements"]),
                   ph_type_map=config["dataset"]["ph_type_map"],
                   processed_data_dir=config["dataset"]["processed_data_dir"], n_pockets_to_sample=ev["n_pockets"],
                   pharms_per_pocket=ev["pharms_per_pocket"], sample_interval=ev["sample_interval"],
                   val_loss_interval=ev["val_loss_interval"], batch_size=config["training"]["batch_size"],
                   graph_config=config["graph"], dynamics_config=config["dynamics"],
                   lr_scheduler_config=config["lr_scheduler"], **config["diffusion"])

    @classmethod
    def load_from_checkpoint(cls, path, map_location="cpu", **overrides) -> "PharmacophoreDiff":
        """Reads a Lightning checkpoint dict written by the reference's train.py (hyper_parameters + state_dict)."""
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        hp = dict(ckpt["hyper_parameters"])
        hp.update(overrides)
        model = cls(**hp)
        model.load_state_dict(ckpt["state_dict"], strict=True)
        return model

    def save_checkpoint(self, path, **extra):
        """A Lightning-format `.ckpt` (checkpoint.save_lightning_checkpoint): loadable by the reference's
        `PharmacophoreDiff.load_from_checkpoint` and by this class."""
        from .checkpoint import save_lightning_checkpoint
        save_lightning_checkpoint(self, path, **extra)

    @property
    def device(self):
        return next(self.dynamics.parameters()).device

    def make_batch(self, pockets: Sequence[Pocket], n_pharms: Sequence[Sequence[int]], device=None,
                   graph_range: Optional[range] = None) -> GraphBatch:
        """copy_graph + dgl.batch of the reference (generate_pharmacophores.py:333-334)."""
        if device is None:   # the reference moves the batch to self.device (pharmacodiff.py:566); the kernels need a GPU, so
            # a model whose parameters still sit on the host samples on the current CUDA device
            device = self.device if self.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        return GraphBatch.from_pockets(pockets, n_pharms, device,
                                       pp_cutoff=self.graph_cutoffs.get("pp", 3.5), pf_k=self.dynamics.pf_k,
                                       graph_range=graph_range)

    # ------------------------------------------------------------------ schedule (pharmacodiff.py:140-160)
    def sigma(self, gamma):
        return torch.sqrt(torch.sigmoid(gamma))

    def alpha(self, gamma):
        return torch.sqrt(torch.sigmoid(-gamma))

    def sigma_and_alpha_t_given_s(self, gamma_t, gamma_s):
        sigma2_t_given_s = -torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t))
        log_alpha2_t, log_alpha2_s = F.logsigmoid(-gamma_t), F.logsigmoid(-gamma_s)
        alpha_t_given_s = torch.exp(0.5 * (log_alpha2_t - log_alpha2_s))
        return sigma2_t_given_s, torch.sqrt(sigma2_t_given_s), alpha_t_given_s, torch.exp(0.5 * log_alpha2_s)

    def step_tables(self):
        """Per-step host tables for the sampling loop, in loop order (s = T-1 .. 0): t value and the three
        posterior coefficients of sample_p_zs_given_zt (pharmacodiff.py:387-400), computed with the same fp32
        torch ops on the host: bit-identical to the reference run on the CPU (the CPU oracle); a reference run on a
        CUDA device evaluates softplus / expm1 / logsigmoid with the device's libm and may differ in the last bit."""
        T = self.n_timesteps
        s = torch.arange(T - 1, -1, -1)
        s_arr, t_arr = s.float() / T, (s + 1).float() / T
        gam = self.gamma.gamma.detach().float().cpu()
        g_s, g_t = gam[torch.round(s_arr * T).long()], gam[torch.round(t_arr * T).long()]
        sigma2_ts, sigma_ts, alpha_ts, _ = self.sigma_and_alpha_t_given_s(g_t, g_s)
        sigma_s, sigma_t = self.sigma(g_s), self.sigma(g_t)
        var_terms = sigma2_ts / alpha_ts / sigma_t
        sigma_q = sigma_ts * sigma_s / sigma_t
        # endpoint parameterisation (pharmacodiff.py:413-418): mu = c1 z_t + c2 pred, same operator order as upstream
        _, _, _, alpha_s = self.sigma_and_alpha_t_given_s(g_t, g_s)
        ep_c1 = alpha_ts * (sigma_s ** 2) / (sigma_t ** 2)
        ep_c2 = alpha_s * sigma2_ts / (sigma_t ** 2)
        return [np.ascontiguousarray(v.numpy(), dtype=np.float32)
                for v in (t_arr, alpha_ts, var_terms, sigma_q, ep_c1, ep_c2)]

    @property
    def ep_mode(self) -> int:
        """PF_EP_COORD | PF_EP_FEAT bits of include/pharmacoforge_b200.h."""
        return (1 if self.endpoint_param_coord else 0) | (2 if self.endpoint_param_feat else 0)

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    @on_batch_device
    def sample_given_receptor(self, g: GraphBatch, init_pharm_com: Optional[torch.Tensor] = None,
                              visualize_trajectory: bool = False, noise: Optional[torch.Tensor] = None,
                              n_steps: Optional[int] = None, return_tensors: bool = False):
        """pharmacodiff.py:433-514.  `noise` ([T+1, Nf, 3+nf] on any device, row 0 = z_T, row 1+i = i-th step, x
        columns first) injects the Gaussian draws for parity runs; by default they come from torch's CUDA
        generator in the reference's call order.  `n_steps` < T stops early (teacher-forced tests)."""
        dev, T, nh = g.device, self.n_timesteps, self.n_pharm_feats
        steps = T if n_steps is None else int(n_steps)
        st = self.dynamics.bind(g)
        nfn = g.n_pharm
        if noise is None:
            # throughput path: Gaussian draws come from Philox inside the posterior kernel (no noise tensors, no 2 (T+1)
            # randn launches); the 64-bit key is drawn from torch's CUDA generator, so torch.manual_seed reproduces a run
            nx = nhh = None
            st.noise_seed.copy_(torch.randint(0, 2 ** 62, (1,), device=dev, dtype=torch.int64))
        else:
            noise = noise.to(dev, non_blocking=True).float()
            nx = noise[:steps + 1, :, 0:3].contiguous()
            nhh = noise[:steps + 1, :, 3:3 + nh].contiguous()
        # frame set-up (pharmacodiff.py:442-452)
        init_prot_com = ops.segment_mean3(g.prot_x, g.prot_ptr)
        if init_pharm_com is None:
            init_pharm_com = init_prot_com
        init_pharm_com = init_pharm_com.to(dev).float().contiguous()
        ops.segment_shift3(g.prot_x, g.prot_ptr, init_pharm_com, -1.0)
        if nx is None:   # z_T: x before h (pharmacodiff.py:455-456) = Philox streams 0 and 1 of step 0
            ops.philox_normal(g.pharm_x, st.noise_seed, 0, 0)
            ops.philox_normal(g.pharm_h, st.noise_seed, 1, 0)
        else:
            g.pharm_x.copy_(nx[0])
            g.pharm_h.copy_(nhh[0])
        self._tables = self.step_tables()   # once per call; kept alive while C reads them
        frames = None
        if visualize_trajectory:
            # device-side trajectory buffer instead of a graph copy + D2H per step (pharmacodiff.py:360-378)
            frames = (torch.empty(steps + 1, nfn, 3, device=dev), torch.empty(steps + 1, nfn, nh, device=dev))
            self._record_frame(g, init_prot_com, frames, 0)
            for i in range(steps):
                self._run_steps(g, st, nx, nhh, i, 1)
                self._record_frame(g, init_prot_com, frames, i + 1)
        else:
            self._run_steps(g, st, nx, nhh, 0, steps)
        # final frame restore (pharmacodiff.py:480-488)
        com = ops.segment_mean3(g.prot_x, g.prot_ptr)
        x0 = g.pharm_x.clone()
        ops.segment_shift3(x0, g.pharm_ptr, com, -1.0)
        ops.segment_shift3(x0, g.pharm_ptr, init_prot_com, 1.0)
        ops.segment_shift3(g.prot_x, g.prot_ptr, com, -1.0)
        ops.segment_shift3(g.prot_x, g.prot_ptr, init_prot_com, 1.0)
        h0 = g.pharm_h * self.pharm_feat_norm_constant
        g.check_status()
        self._tables = None   # read on the host while the steps were enqueued; never cached across calls
        if return_tensors:
            return x0, h0
        x0_h, h0_h = x0.cpu(), h0.cpu()
        fr = None
        if frames is not None:
            # frames carry x_t moved back to the input frame and h_t as is: the reference's get_pos_feat_for_visual
            # un-normalises h_0 only (pharmacodiff.py:84-86, 360-378), never the h_t it records
            fr = (frames[0].cpu(), frames[1].cpu())
        ptr = g.pharm_ptr_host
        out = []
        for b in range(g.n_graphs):
            sl = slice(int(ptr[b]), int(ptr[b + 1]))
            tf = (fr[0][:, sl], fr[1][:, sl]) if fr is not None else None
            out.append(SampledPharmacophore(x0_h[sl], h0_h[sl], self.ph_type_map, traj_frames=tf))
        return out

    def _run_steps(self, g, st, nx, nhh, first: int, count: int):
        a = st.args
        T = self.n_timesteps
        off = first  # tables are in loop order; the same offset applies to every table
        if getattr(self, "_tables", None) is None:      # direct callers (tests) that did not go through the sampler
            self._tables = self.step_tables()
        t_host, alpha_ts, var_terms, sigma_q, ep_c1, ep_c2 = self._tables
        a.t_host = t_host[off:].ctypes.data
        a.alpha_ts_host = alpha_ts[off:].ctypes.data
        a.var_terms_host = var_terms[off:].ctypes.data
        a.sigma_q_host = sigma_q[off:].ctypes.data
        a.ep_c1_host = ep_c1[off:].ctypes.data
        a.ep_c2_host = ep_c2[off:].ctypes.data
        a.ep_mode = self.ep_mode
        philox = nx is None
        a.noise_x = nx[1 + first:].data_ptr() if (count and not philox) else 0
        a.noise_h = nhh[1 + first:].data_ptr() if (count and not philox) else 0
        a.noise_seed = st.noise_seed.data_ptr()
        a.noise_step0 = first + 1
        a.n_steps = count
        if not count:
            return
        if not (philox and self.use_cuda_graph):
            ops.sample_loop(g.pharm_x, g.pharm_h, g.prot_x, st.addr)
            st.warmed = True
            return
        # The whole loop as ONE CUDA graph (Philox mode only: every kernel argument is then the same from call to call --
        # buffers of this batch, the schedule constants, the device-resident seed).  Captured on the second use of a
        # (batch, first, count, flags) combination: the first one runs eagerly so that every kernel has been configured.
        key = (first, count, int(a.flags), int(a.ep_mode))
        gr = st.graphs.get(key)
        if gr is None and not st.warmed:
            ops.sample_loop(g.pharm_x, g.pharm_h, g.prot_x, st.addr)
            st.warmed = True
            return
        if gr is None:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                ops.sample_loop(g.pharm_x, g.pharm_h, g.prot_x, st.addr)
            st.graphs[key] = gr
        gr.replay()

    def _record_frame(self, g, init_prot_com, frames, i):
        prot_com = ops.segment_mean3(g.prot_x, g.prot_ptr)
        x = g.pharm_x.clone()
        ops.segment_shift3(x, g.pharm_ptr, init_prot_com - prot_com, 1.0)
        frames[0][i] = x
        frames[1][i] = g.pharm_h

    def sample(self, ref_graphs: Sequence[Pocket], n_pharms: List[List[int]], max_batch_size: int = 32,
               init_pharm_com: Optional[torch.Tensor] = None, visualize_trajectory: bool = False,
               noise: Optional[torch.Tensor] = None, n_steps: Optional[int] = None):
        """pharmacodiff.py:516-578: pockets x samples flattened pocket-major, chunked by max_batch_size, regrouped per
        pocket.  `init_pharm_com` [n_pockets, 3] defaults to each pocket's mean position (:531-535) and is indexed by the
        pocket of every graph of a chunk (:563).  `noise` ([T+1, Nf_total, 3+nf], columns = pharmacophore nodes in the
        flattened graph order) and `n_steps` are the parity hooks of `sample_given_receptor`."""
        if init_pharm_com is None:
            init_pharm_com = torch.stack([r.prot_x.float().mean(dim=0) for r in ref_graphs])
        flat_ref = [r for r, szs in enumerate(n_pharms) for _ in szs]
        flat_nf = [int(n) for szs in n_pharms for n in szs]
        n_total = len(flat_ref)
        sampled: List[SampledPharmacophore] = []
        col = 0
        for b in range(ceil(n_total / max_batch_size)):
            rng = range(b * max_batch_size, min((b + 1) * max_batch_size, n_total))
            g = self.make_batch(ref_graphs, n_pharms, graph_range=rng)
            coms = init_pharm_com[flat_ref[rng.start:rng.stop]]
            nf = sum(flat_nf[rng.start:rng.stop])
            chunk_noise = None if noise is None else noise[:, col:col + nf]
            col += nf
            sampled.extend(self.sample_given_receptor(g, init_pharm_com=coms, visualize_trajectory=visualize_trajectory,
                                                      noise=chunk_noise, n_steps=n_steps))
        out, end = [], 0
        for szs in n_pharms:
            out.append(sampled[end:end + len(szs)])
            end += len(szs)
        return out

    @on_batch_device
    def forward(self, g: GraphBatch, phase: str = "train", t_int: Optional[torch.Tensor] = None,
                eps: Optional[Dict[str, torch.Tensor]] = None):
        """The training / validation objective of pharmacodiff.py:162-243 (eps parameterisation): returns
        (losses, metrics) with the reference's keys.  `g` carries the ground truth set by
        `GraphBatch.set_pharmacophores(x_0, h_0)`; `t_int` ([B] integer timesteps in [0, T)) and `eps` ({'x': [Nf,3],
        'h': [Nf,F]}) inject the reference's `torch.randint` / `torch.randn` draws (h is drawn before x,
        pharmacodiff.py:189-192) for parity runs.

        In training mode with gradients enabled the eps prediction runs through the differentiable custom ops of
        `train_ops.py` (`train_graph.dynamics_forward`: hand-written forward + backward kernels, training-mode dropout),
        so the returned losses carry an autograd graph onto the reference-named parameters (which must live on the GPU).
        Otherwise (eval / no_grad) it runs the fused sampling kernels."""
        if getattr(g, "pharm_x0", None) is None:
            raise ValueError("forward() needs the ground-truth pharmacophores: call g.set_pharmacophores(x_0, h_0)")
        dev, T = g.device, self.n_timesteps
        B, fb = g.n_graphs, g.batch_idxs()["pharm"]
        differentiable = self.training and torch.is_grad_enabled()
        with torch.no_grad():
            if t_int is None:
                t_int = torch.randint(0, T, size=(B,), device=dev)
            t = t_int.to(dev).float() / T
            if eps is None:
                eps = {"h": torch.randn(g.n_pharm, self.n_pharm_feats, device=dev),
                       "x": torch.randn(g.n_pharm, 3, device=dev)}
            eps_x, eps_h = eps["x"].to(dev).float(), eps["h"].to(dev).float()
            # The reference shifts pharm x_0 and prot x_0 together and in place; here the ground truth stays untouched
            # in g.pharm_x0, so the protein must start from its input frame too -- otherwise a second forward() on the
            # same batch (cached validation batches, several epochs) would subtract the pharmacophore COM twice.
            g.prot_x.copy_(g.prot_x0)
            h0 = g.pharm_h0 / self.pharm_feat_norm_constant                   # normalize, :81-83
            x0 = g.pharm_x0.clone()
            com0 = ops.segment_mean3(x0, g.pharm_ptr)                          # com_removal(pharm_feat='x_0'), :178
            ops.segment_shift3(x0, g.pharm_ptr, com0, -1.0)
            ops.segment_shift3(g.prot_x, g.prot_ptr, com0, -1.0)
            gamma_t = self.gamma(t.to(self.gamma.gamma.device)).to(dev)
            alpha_t = self.alpha(gamma_t)[fb][:, None]
            sigma_t = self.sigma(gamma_t)[fb][:, None]
            if g.pharm_h is None or g.pharm_h.shape[1] != self.n_pharm_feats:
                g.pharm_h = torch.zeros(max(g.n_pharm, 1), self.n_pharm_feats, device=dev)
            g.pharm_x.copy_(alpha_t * x0 + sigma_t * eps_x)                    # noised_representation, :110-127
            g.pharm_h.copy_(alpha_t * h0 + sigma_t * eps_h)
            com_t = None
            if self.remove_com:                                                # :123-125
                com_t = ops.segment_mean3(g.pharm_x, g.pharm_ptr)
                ops.segment_shift3(g.pharm_x, g.pharm_ptr, com_t, -1.0)
                ops.segment_shift3(g.prot_x, g.prot_ptr, com_t, -1.0)
        if differentiable:
            if not next(self.dynamics.parameters()).is_cuda:
                raise RuntimeError("training needs the model parameters on the GPU: call model.to(g.device)")
            from . import train_graph
            h_dyn, x_dyn = train_graph.dynamics_forward(self.dynamics, g, t, training=True)
        else:
            with torch.no_grad():
                self.dynamics.check_status_every_call = False
                try:
                    h_dyn, x_dyn = self.dynamics(g, t)
                finally:
                    self.dynamics.check_status_every_call = True
        g.check_status()
        if self.endpoint_param_feat:       # the network predicts h_0: cross entropy on its logits (pharmacodiff.py:204-206)
            h0_pred = h_dyn
            h_loss = F.cross_entropy(h0_pred, h0.argmax(dim=1), reduction="none")
        else:
            h_loss = (eps_h - h_dyn).square().sum(dim=1)
            h0_pred = (g.pharm_h - sigma_t * h_dyn) / alpha_t
        if self.endpoint_param_coord:      # ... and x_0, in the frame before the COM of x_t was removed (:210-216)
            x0_pred = x_dyn + com_t[fb] if self.remove_com else x_dyn
            x_loss = (x0_pred - x0).square().sum(dim=1)
        else:
            x_loss = (eps_x - x_dyn).square().sum(dim=1)
            x0_pred = (g.pharm_x - sigma_t * x_dyn) / alpha_t
        w_metric = 1 - t[fb]
        w_loss = w_metric if self.weighted_loss else torch.ones_like(w_metric)
        losses = {phase + " pos loss": (x_loss * w_loss).sum() / eps_x.numel(),
                  phase + " feat loss": (h_loss * w_loss).sum() / eps_h.numel()}
        with torch.no_grad():
            sq = (x0_pred - x0).square().sum(dim=1)
            hit = (h0_pred.argmax(dim=1) == h0.argmax(dim=1)).float()
            metrics = {phase + " position error": sq.mean(), phase + " weighted position error": (w_metric * sq).mean(),
                       phase + " accuracy": hit.mean(), phase + " weighted accuracy": (w_metric * hit).mean()}
        return losses, metrics

    def validation_step(self, g: GraphBatch, batch_idx: int = 0, **inject):
        """pharmacodiff.py:299-318 without the Lightning logging: total loss / total error of one batch."""
        phase = "val"
        losses, metrics = self.forward(g, phase=phase, **inject)
        losses[phase + " total loss"] = torch.stack(list(losses.values())).sum()
        metrics[phase + " total error"] = metrics[phase + " position error"] + 1 - metrics[phase + " accuracy"]
        metrics[phase + " weighted total error"] = (metrics[phase + " weighted position error"] + 1 -
                                                    metrics[phase + " weighted accuracy"])
        return losses, metrics

    def training_step(self, g: GraphBatch, batch_idx: int = 0, **inject):
        """pharmacodiff.py:265-297 without the Lightning logging / periodic sampling: the total loss (pos + feat) of one
        batch, carrying the autograd graph (call `.backward()` on it), plus the loss and metric dicts."""
        phase = "train"
        losses, metrics = self.forward(g, phase=phase, **inject)
        losses[phase + " total loss"] = torch.stack(list(losses.values())).sum()
        metrics[phase + " total error"] = metrics[phase + " position error"] + 1 - metrics[phase + " accuracy"]
        metrics[phase + " weighted total error"] = (metrics[phase + " weighted position error"] + 1 -
                                                    metrics[phase + " weighted accuracy"])
        return losses[phase + " total loss"], losses, metrics

    def configure_optimizers(self):
        """pharmacodiff.py:254-263: Adam(base_lr, weight_decay) + ReduceLROnPlateau from lr_scheduler_config."""
        cfg = self.lr_scheduler_config or {}
        opt = torch.optim.Adam(self.parameters(), lr=cfg.get("base_lr", 1e-4), weight_decay=cfg.get("weight_decay", 0.0))
        # The packed kernel weights are cached on (data_ptr, _version) of every parameter.  Not every in-place update
        # bumps the version counter (measured: torch.optim.Adam(fused=True) does not), so a step also drops the cache.
        opt.register_step_post_hook(lambda *_: setattr(self.dynamics, "_packed_key", None))
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, **cfg.get("reducelronplateau", {}))
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "monitor": cfg.get("monitor"),
                                                   "interval": cfg.get("interval"), "frequency": cfg.get("frequency")}}
