"""Host-only helpers (numpy / torch CPU): nothing here touches the CUDA library, so tools that only need the noise
schedule or synthetic inputs -- bench.py's reference arm, the oracle checks -- can import them without loading
libpharmacoforge_b200.so."""
from __future__ import annotations

import numpy as np
import torch


def polynomial_gamma(timesteps: int, precision: float, power: float) -> torch.Tensor:
    """gamma_t = -(log alpha_t^2 - log sigma_t^2) of the clipped polynomial schedule, float64 on the host then
    float32, as the reference builds it once at construction (pharmacodiff.py:602-664)."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    a2 = (1.0 - np.power(x / steps, power)) ** 2
    step = np.clip(np.concatenate([np.ones(1), a2])[1:] / np.concatenate([np.ones(1), a2])[:-1], 0.001, 1.0)
    a2 = (1.0 - 2.0 * precision) * np.cumprod(step) + precision
    return torch.from_numpy(-(np.log(a2) - np.log(1.0 - a2))).float()
