#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4]): one optimisation step = forward + backward + gradient
all-reduce + Adam on a batch of 64 distinct synthetic pockets per GPU (N ~ U(250, 600) atoms, one pharmacophore of 4..8
centres each, dev.yml model, dropout 0.1), through PharmacophoreDiff.training_step.

    python bench_train.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_train.py --gpus N ...

Not the headline benchmark (that is bench.py, sampling on configs[1]); this first training path is unfused.  One JSON
line on rank 0; `cpu_baseline` times the same step through the oracle's autograd on the host cores (bounded sample).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from bench import CUT, DYN, PH_TYPES, T_STEPS, ClockSampler, load_weights


def make_batch_inputs(n_pockets, seed0):
    from pharmacoforge_b200.batch import Pocket
    from pharmacoforge_b200.synthetic import make_pocket
    rng = np.random.default_rng(seed0)
    pockets, sizes, x0, types = [], [], [], []
    for i in range(n_pockets):
        n = int(rng.integers(250, 601))
        pos, onehot = make_pocket(n, seed=seed0 * 1000 + i)
        pockets.append(Pocket.from_numpy(pos, onehot))
        nf = int(rng.integers(4, 9))
        sizes.append([nf])
        x0.append(pos.mean(0, keepdims=True) + rng.normal(size=(nf, 3)) * 3.0)
        types.append(rng.integers(0, 6, size=nf))
    x0 = torch.from_numpy(np.concatenate(x0).astype(np.float32))
    h0 = torch.nn.functional.one_hot(torch.from_numpy(np.concatenate(types)), 6).float()
    return pockets, sizes, x0, h0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="pockets (graphs) per GPU per step")
    ap.add_argument("--cpu-pockets", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")   # stdout carries exactly one JSON line; other writers of fd 1 go to stderr
    os.dup2(2, 1)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench_train.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pharmacoforge_b200 import _lib
    from pharmacoforge_b200.batch import GraphBatch
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    from pharmacoforge_b200.sharding import allreduce_gradients
    lib = _lib.load()
    sd = load_weights()
    model = PharmacophoreDiff(6, 11, PH_TYPES, n_timesteps=T_STEPS, graph_config={"graph_cutoffs": CUT},
                              dynamics_config=DYN, precision=1e-5,
                              lr_scheduler_config={"base_lr": 1e-4, "weight_decay": 1e-12})
    model.load_state_dict(sd)
    model = model.to(dev).train()
    opt = model.configure_optimizers()["optimizer"]
    pockets, sizes, x0, h0 = make_batch_inputs(args.batch, seed0=1 + rank)
    h2d = 0

    def step():
        nonlocal h2d
        g = GraphBatch.from_pockets(pockets, sizes, dev)          # host pockets -> device batch + pp graph (every step,
        g.set_pharmacophores(x0, h0)                              # as a data loader would hand over a new batch)
        h2d = g.h2d_bytes + x0.numel() * 4 + h0.numel() * 4
        opt.zero_grad(set_to_none=True)
        total, _, _ = model.training_step(g)
        total.backward()
        n = allreduce_gradients(model)
        opt.step()
        return total, n, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = lib.pf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses = []
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            total, nflat, g = step()
            losses.append(float(total))                            # device -> host read of the step's loss
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = lib.pf_launch_count() - l0
    # opt-in, reported separately: the last layer's protein side (never read, never visited by autograd) is not run
    model.dynamics.skip_dead_work = True
    for _ in range(2):
        step()
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(args.steps):
        total, _, _ = step()
        float(total)
    d1.record()
    barrier()
    ms_dead = torch.tensor([d0.elapsed_time(d1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_dead, op=dist.ReduceOp.MAX)
    ms_dead = float(ms_dead.item())
    model.dynamics.skip_dead_work = False
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pf_oracle as O
            torch.set_num_threads(os.cpu_count())
            cp, cs, cx0, ch0 = make_batch_inputs(args.cpu_pockets, seed0=1)
            b = O.build_batch([(p.prot_x, p.prot_h) for p in cp], cs)
            sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and k != "gamma.gamma" and v.numel() else v)
                   for k, v in sd.items()}
            nf = cx0.shape[0]
            t0 = time.perf_counter()
            lo, _ = O.forward_loss(sdg, b, cx0, ch0, torch.randint(0, T_STEPS, (args.cpu_pockets,)), torch.randn(nf, 3),
                                   torch.randn(nf, 6), T_STEPS, sd["gamma.gamma"], dict(DYN, graph_cutoffs=CUT))
            torch.stack(list(lo.values())).sum().backward()
            dt = time.perf_counter() - t0
            cpu = {"value": args.cpu_pockets / dt, "unit": "graphs/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"forward + backward of {args.cpu_pockets} pockets through the oracle's autograd in {dt:.2f} s"}
        value = world * args.batch * args.steps / (ms / 1e3)
        print(json.dumps({
            "metric": "training graphs/sec (forward + backward + all-reduce + Adam)", "value": value, "unit": "graphs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4]: {args.batch} distinct synthetic pockets per GPU per step, N~U(250,600) atoms, "
                                   "pharmacophore sizes 4..8, dev.yml model, dropout 0.1, Adam",
                       "prot_nodes": g.n_prot, "pp_edges": g.n_pp_edges, "flat_gradient_elements": nflat,
                       "parallelism": f"data parallel x{world}, one flat fp32 all-reduce per step"},
            "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "api": "GraphBatch.from_pockets + training_step + backward + allreduce_gradients + Adam.step"},
            "final_loss": losses[-1], "first_loss": losses[0], "gpu_launches": int(launches), "cpu_baseline": cpu,
            "exact_dead_work_elimination": {
                "value": world * args.batch * args.steps / (ms_dead / 1e3), "unit": "graphs/s",
                "ms_per_step": ms_dead / args.steps,
                "note": "NOT the headline: dynamics.skip_dead_work -- the last conv layer's pp / fp messages and protein "
                        "update (never read, dynamics_gvp.py:84-92; autograd never visits them) are not run; same losses "
                        "and gradients (tests/test_gpu_training.py)"},
            "clocks": clk.summary(), "note": "kernel-per-op training path (per-edge tensors are materialised; every arithmetic "
                                             "node is a hand-written CUDA kernel, train_ops.py), one autograd node per "
                                             "message chain / node update (train_fused.py)"}), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
