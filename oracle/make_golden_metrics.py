"""Generate tests/golden/complementarity.npz with the reference's own compute_complementarity (analysis/metrics.py:54-86)
on seeded random inputs.  Test infrastructure only.      python oracle/make_golden_metrics.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_loader  # noqa: E402

reference_loader.load()
from pharmacoforge.analysis.metrics import compute_complementarity  # noqa: E402
from pharmacoforge.constants import ph_idx_to_type  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
gen = torch.Generator().manual_seed(11)
out = {}
for case in range(6):
    n, m = int(torch.randint(3, 9, (1,), generator=gen)), int(torch.randint(5, 40, (1,), generator=gen))
    ppos = torch.randn(n, 3, generator=gen) * 4
    rpos = torch.randn(m, 3, generator=gen) * 5
    pt = torch.randint(0, 6, (n,), generator=gen)
    rt = torch.randint(0, 6, (m,), generator=gen)
    cnt = compute_complementarity([ph_idx_to_type[int(i)] for i in pt], ppos, [ph_idx_to_type[int(i)] for i in rt], rpos,
                                  return_count=True)
    out.update({f"ppos{case}": ppos.numpy(), f"rpos{case}": rpos.numpy(), f"pt{case}": pt.numpy(), f"rt{case}": rt.numpy(),
                f"count{case}": np.asarray(int(cnt))})
np.savez(os.path.join(GOLD, "complementarity.npz"), n_cases=6, **out)
print({k: int(v) for k, v in out.items() if k.startswith("count")})
