import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import CUT, DYN, PH_TYPES, T_STEPS, load_weights
from bench_train import make_batch_inputs
from pharmacoforge_b200.batch import GraphBatch
from pharmacoforge_b200.diffusion import PharmacophoreDiff
from pharmacoforge_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
sd = load_weights()
model = PharmacophoreDiff(6, 11, PH_TYPES, n_timesteps=T_STEPS, graph_config={"graph_cutoffs": CUT}, dynamics_config=DYN, precision=1e-5, lr_scheduler_config={"base_lr": 1e-4})
model.load_state_dict(sd); model = model.to(dev).train()
opt = model.configure_optimizers()["optimizer"]
pockets, sizes, x0, h0 = make_batch_inputs(64, 1)
def tm(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
acc = {}
for it in range(6):
    g, a = tm(lambda: GraphBatch.from_pockets(pockets, sizes, dev).set_pharmacophores(x0, h0))
    opt.zero_grad(set_to_none=True)
    l0 = lib.pf_launch_count()
    (total, _, _), b = tm(lambda: model.training_step(g))
    l1 = lib.pf_launch_count()
    _, c = tm(lambda: total.backward())
    l2 = lib.pf_launch_count()
    _, d = tm(lambda: opt.step())
    if it >= 2:
        for k, v in (("batch", a), ("forward", b), ("backward", c), ("adam", d)): acc[k] = acc.get(k, 0) + v / 4
print({k: round(v, 2) for k, v in acc.items()}, "launches fwd", l1 - l0, "bwd", l2 - l1)
from torch.profiler import profile, ProfilerActivity
g = GraphBatch.from_pockets(pockets, sizes, dev).set_pharmacophores(x0, h0)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    total, _, _ = model.training_step(g); total.backward(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
