import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def layout():
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return load


@pytest.fixture(scope="session")
def sd(layout):
    """The weights the golden fixtures were generated with (deterministic per key)."""
    import pf_oracle
    from pharmacoforge_b200.synthetic import synth_state_dict
    out = synth_state_dict(layout, seed=0)
    out["gamma.gamma"] = pf_oracle.gamma_table(100, 1e-5)
    return out


DEV_DYNAMICS = dict(vector_size=16, n_convs=2, n_hidden_scalars=128, message_norm="mean", dropout=0.1, ff_k=0, pf_k=5,
                    n_message_gvps=3, n_update_gvps=2, n_noise_gvps=4,
                    graph_cutoffs={"pp": 3.5, "pf": 8, "fp": 8, "ff": 9})


@pytest.fixture(scope="session")
def dyn_cfg():
    return dict(DEV_DYNAMICS)
