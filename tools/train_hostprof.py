"""Host-side profile of one training step (cProfile, cumulative + self time) and the torch profiler table sorted by
self CPU time.  usage: train_hostprof.py"""
import sys, time, os, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import CUT, DYN, PH_TYPES, T_STEPS, load_weights
from bench_train import make_batch_inputs
from pharmacoforge_b200.batch import GraphBatch
from pharmacoforge_b200.diffusion import PharmacophoreDiff
dev = torch.device("cuda:0")
sd = load_weights()
model = PharmacophoreDiff(6, 11, PH_TYPES, n_timesteps=T_STEPS, graph_config={"graph_cutoffs": CUT}, dynamics_config=DYN, precision=1e-5, lr_scheduler_config={"base_lr": 1e-4})
model.load_state_dict(sd); model = model.to(dev).train()
opt = model.configure_optimizers()["optimizer"]
pockets, sizes, x0, h0 = make_batch_inputs(64, 1)
def step():
    g = GraphBatch.from_pockets(pockets, sizes, dev).set_pharmacophores(x0, h0)
    opt.zero_grad(set_to_none=True)
    total, _, _ = model.training_step(g)
    total.backward()
    opt.step()
    return float(total)
for _ in range(4): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable()
for key in ("tottime", "cumtime"):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28); print(s.getvalue()[:6000])
