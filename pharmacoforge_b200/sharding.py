"""Sharding of independent graphs across GPUs.

Sampling graphs never interact (every edge builder and reduction is restricted to one graph), so the flattened
pocket-major graph list is cut into `world` contiguous ranges balanced by protein-atom count; each rank runs
its own reverse-diffusion loop with replicated weights and no collective on the path.  Results are gathered
once at the end (`gather_results`).

Training is data parallel: every rank runs forward + backward on its share of the batch and `allreduce_gradients`
averages the gradients with ONE all-reduce of a flat fp32 buffer (NCCL over NVLink on the GPU box), the semantics of
Lightning's DDP strategy the reference trains under (mean of the per-rank gradients; each rank's loss is normalised by
its LOCAL element count, pharmacodiff.py:231-232).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch


def shard_ranges(n_pharms: Sequence[Sequence[int]], world: int,
                 pocket_atoms: Optional[Sequence[float]] = None) -> List[range]:
    """Contiguous graph ranges [start, stop) per rank over the flattened (pocket, sample) list, balanced by the weight
    of each graph: `pocket_atoms[p]` is any per-pocket cost proxy -- the atom count, or better the pocket's pp edge
    count (the edge kernels are ~80 % of a step), which is what bench.py's configs[3] workload passes."""
    counts = np.fromiter((len(s) for s in n_pharms), dtype=np.int64, count=len(n_pharms))
    w = np.asarray(pocket_atoms, dtype=np.float64) if pocket_atoms is not None else np.ones(len(n_pharms))
    flat_w = np.repeat(w, counts)
    n = int(flat_w.size)
    if n == 0:
        return [range(0, 0) for _ in range(world)]
    cum = np.concatenate([[0.0], np.cumsum(flat_w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        cuts.append(int(np.clip(np.searchsorted(cum, target, side="left"), cuts[-1], n)))
    cuts.append(n)
    return [range(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_results(local: torch.Tensor, group=None) -> Optional[List[torch.Tensor]]:
    """The one collective of the sampling path: rank 0 receives every rank's [n_local, C] result rows."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if dist.get_backend(group) == "gloo" and local.is_cuda:
        # gloo moves host memory: stage the (small) result rows through the CPU; NCCL gathers device to device
        parts = gather_results(local.cpu(), group)
        return None if parts is None else [p.to(local.device) for p in parts]
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    n_max = int(max(int(c.item()) for c in counts))
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return [b[:int(c.item())] for b, c in zip(bufs, counts)]


def allreduce_gradients(module: torch.nn.Module, group=None) -> int:
    """Average the gradients of `module` over the ranks with a single all-reduce of one flat fp32 buffer.

    Every rank lays out ALL parameters in `named_parameters()` order; parameters without a gradient (the protein side of
    the last conv layer is never used, so autograd leaves `grad=None` -- the same set on every rank) contribute zeros and
    keep `grad=None` afterwards, so Adam skips them exactly as it does in the reference.  Returns the buffer length."""
    import torch.distributed as dist
    params = [p for _, p in module.named_parameters() if p.requires_grad and p.numel() > 0]
    if not params:
        return 0
    dev = next((p.grad.device for p in params if p.grad is not None), params[0].device)
    # one concatenation in, one all-reduce, one multi-tensor copy out (no per-parameter kernel launches)
    zmax = max((p.numel() for p in params if p.grad is None), default=0)
    zeros = torch.zeros(zmax, dtype=torch.float32, device=dev)
    flat = torch.cat([p.grad.reshape(-1).to(torch.float32) if p.grad is not None else zeros[:p.numel()] for p in params])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= dist.get_world_size(group)
        live = [p for p in params if p.grad is not None]
        chunks = flat.split([p.numel() for p in params])
        src = [c.view_as(p.grad) for c, p in zip(chunks, params) if p.grad is not None]
        torch._foreach_copy_([p.grad for p in live], src)
    return flat.numel()
