"""Synthetic pockets and pharmacophore-size lists for tests and benchmarks.

There is no dataset and no network, so every measured configuration runs on
pockets drawn here: N heavy atoms rejection-sampled uniformly in a spherical
shell 4 A < r < R around the (empty) binding site, R chosen for a density of
0.055 atoms/A^3, minimum pair distance 1.3 A, element one-hot over the 11
`prot_elements` of configs/dev.yml, then shifted by a random offset so that
centre-of-mass handling is exercised.  numpy only; deterministic per seed.
"""
from __future__ import annotations

import numpy as np

N_PROT_ELEMENTS = 11
_ELEMENT_P = np.array([0.62, 0.17, 0.19, 0.02] + [0.0] * 7)
README_SIZES = [3, 4, 5, 6, 7, 8]


def pocket_radius(n_atoms: int, density: float = 0.055, r_in: float = 4.0) -> float:
    return float((3.0 * n_atoms / (4.0 * np.pi * density) + r_in ** 3) ** (1.0 / 3.0))


def make_pocket(n_atoms: int = 400, seed: int = 0, r_in: float = 4.0, min_dist: float = 1.3,
                offset_range: float = 50.0):
    """Returns (pos [N,3] float32, onehot [N,11] float32)."""
    rng = np.random.default_rng(seed)
    r_out = pocket_radius(n_atoms, r_in=r_in)
    pts = np.empty((n_atoms, 3), dtype=np.float64)
    n = 0
    md2 = min_dist * min_dist
    while n < n_atoms:
        cand = rng.uniform(-r_out, r_out, size=(4 * n_atoms, 3))
        rr = np.sqrt((cand * cand).sum(1))
        cand = cand[(rr > r_in) & (rr < r_out)]
        for c in cand:
            if n and ((pts[:n] - c) ** 2).sum(1).min() < md2:
                continue
            pts[n] = c
            n += 1
            if n == n_atoms:
                break
    types = rng.choice(N_PROT_ELEMENTS, size=n_atoms, p=_ELEMENT_P)
    offset = rng.uniform(-offset_range, offset_range, size=(1, 3))
    pos = (pts + offset).astype(np.float32)
    onehot = np.zeros((n_atoms, N_PROT_ELEMENTS), dtype=np.float32)
    onehot[np.arange(n_atoms), types] = 1.0
    return pos, onehot


def readme_sizes(n_samples: int = 30):
    """[3,4,5,6,7,8] repeated, as in the reference README example (README.md:27)."""
    reps = (n_samples + len(README_SIZES) - 1) // len(README_SIZES)
    return (README_SIZES * reps)[:n_samples]


def uniform_sizes(n_samples: int, lo: int, hi: int, seed: int = 0):
    rng = np.random.default_rng(10_000 + seed)
    return [int(v) for v in rng.integers(lo, hi + 1, size=n_samples)]


def synth_state_dict(layout: dict, seed: int = 0):
    """Deterministic random weights for a {state_dict key: shape} layout.

    The reference's own random init depends on PYTHONHASHSEED (module creation
    order follows a `set`, gvp.py:367,421), so "same seed" does not give the
    same weights twice.  Values here depend only on (seed, key): Linear / Wh / Wu
    entries are U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like the reference's init,
    LayerNorm affine terms are perturbed away from (1, 0) so that they are
    exercised.  `gamma.gamma` is skipped (it is the noise schedule, not a weight).
    """
    import zlib

    import torch

    out = {}
    for key in sorted(layout):
        shape = tuple(int(s) for s in layout[key])
        if key == "gamma.gamma":
            continue
        rng = np.random.default_rng([seed, zlib.crc32(key.encode())])
        if len(shape) == 1 and shape[0] == 0:
            val = np.zeros(shape, dtype=np.float32)
        elif key.endswith("feat_norm.weight") or key.endswith("encoder.2.weight"):
            val = 1.0 + 0.1 * rng.standard_normal(shape)
        elif key.endswith("feat_norm.bias") or key.endswith("encoder.2.bias"):
            val = 0.1 * rng.standard_normal(shape)
        elif key.endswith(".Wh") or key.endswith(".Wu"):
            k = 1.0 / np.sqrt(shape[0])
            val = rng.uniform(-k, k, size=shape)
        elif key.endswith(".weight"):
            k = 1.0 / np.sqrt(shape[1])
            val = rng.uniform(-k, k, size=shape)
        elif key.endswith(".bias"):
            # fan_in of the matching weight is not known from the bias alone; 128-wide layers dominate
            val = rng.uniform(-0.08, 0.08, size=shape)
        else:
            raise KeyError(f"no synthetic rule for {key}")
        out[key] = torch.from_numpy(np.asarray(val, dtype=np.float32))
    return out
