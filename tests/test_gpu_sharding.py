"""Multi-process sampling on the GPU box: the graphs of a job sharded over 2 ranks (shard_ranges, no collective on the
path, one final gather) give, bit for bit, what one process computes for the whole job."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DYN = dict(vector_size=16, n_convs=2, n_hidden_scalars=128, message_norm="mean", dropout=0.1, ff_k=0, pf_k=5,
           n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4)
CUT = {"pp": 3.5, "pf": 8, "fp": 8, "ff": 9}
SPECS = [(400, 0), (250, 1), (120, 2), (333, 3)]
SIZES = [[3, 4, 5, 6, 7, 8], [4, 8], [3, 5, 7], [8, 8, 3, 4]]
STEPS = 7


def _job(dev, sd, graph_range, noise, precision):
    from pharmacoforge_b200.batch import Pocket
    from pharmacoforge_b200.diffusion import PharmacophoreDiff
    from pharmacoforge_b200.synthetic import make_pocket
    model = PharmacophoreDiff(6, 11, list("abcdef"), n_timesteps=100, graph_config={"graph_cutoffs": CUT},
                              dynamics_config=DYN, precision=1e-5)
    model.load_state_dict(sd)
    model.eval()
    model.dynamics.edge_mlp_precision = precision
    pockets = [Pocket.from_numpy(*make_pocket(n, seed=s)) for n, s in SPECS]
    flat = [n for szs in SIZES for n in szs]
    off = np.concatenate([[0], np.cumsum(flat)])
    if len(graph_range) == 0:
        return torch.zeros(0, 9, device=dev)
    g = model.make_batch(pockets, SIZES, device=dev, graph_range=graph_range)
    cols = slice(int(off[graph_range.start]), int(off[graph_range.stop]))
    x, h = model.sample_given_receptor(g, noise=noise[:, cols], n_steps=STEPS, return_tensors=True)
    return torch.cat([x, h], dim=1)


def _worker(rank, world, port, sd, noise, precision, out_dir):
    import torch.distributed as dist
    from pharmacoforge_b200.sharding import gather_results, shard_ranges
    n_dev = torch.cuda.device_count()
    dev = torch.device("cuda", rank % n_dev)
    torch.cuda.set_device(dev)
    backend = "nccl" if n_dev >= world else "gloo"        # one GPU shared by both ranks: gloo carries the final gather
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = shard_ranges(SIZES, world, pocket_atoms=[n for n, _ in SPECS])[rank]
        parts = gather_results(_job(dev, sd, rng, noise, precision))
        if rank == 0:
            torch.save({"rows": torch.cat(parts).cpu(), "counts": [p.shape[0] for p in parts], "backend": backend},
                       os.path.join(out_dir, "sharded.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_two_rank_sharded_sampling_is_bit_identical(tmp_path, sd, precision):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    nf = sum(n for szs in SIZES for n in szs)
    noise = torch.randn(STEPS + 1, nf, 9, generator=torch.Generator().manual_seed(31))
    sd_cpu = {k: v.clone() for k, v in sd.items()}
    mp.spawn(_worker, args=(2, port, sd_cpu, noise, precision, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "sharded.pt"))
    n_graphs = sum(len(s_) for s_ in SIZES)
    whole = _job(torch.device("cuda:0"), sd_cpu, range(0, n_graphs), noise, precision).cpu()
    assert sum(got["counts"]) == nf and min(got["counts"]) > 0          # both ranks held work
    assert torch.equal(got["rows"], whole), f"sharded ({got['backend']}) != single process"
