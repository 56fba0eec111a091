"""Golden fixture for the checkpoint / config I/O, from the reference's OWN constructor.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference).  The reference model is built by its own
`model_from_config(configs/dev.yml)` over oracle/shims; what its constructor hands to Lightning's `save_hyperparameters()`
(pharmacodiff.py:78) -- i.e. the `hyper_parameters` entry of every checkpoint train.py writes -- is stored together
with the parsed config, so that the tests can assemble a Lightning-format `.ckpt` (state_dict from the deterministic
`synth_state_dict`) and a `run_dir/config.yaml` exactly as generate_pharmacophores.py:231-269 expects to find them.

    python oracle/make_golden_ckpt.py          # writes tests/golden/lightning_ckpt_meta.json
"""
import json
import os
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import reference_loader  # noqa: E402
from make_golden import GOLD, build_model  # noqa: E402


def main():
    model, cfg, layout = build_model()
    hp = dict(model.hparams)
    json.dumps(hp)   # must be plain data (it is: numbers, strings, lists, dicts)
    # the reference's multi-pharmacophore xyz writer on a fixed input (utils/unorganized_utils.py:111-128)
    import torch
    from pharmacoforge.utils import write_pharmacophore_file
    gen = torch.Generator().manual_seed(5)
    coords = [torch.randn(3, 3, generator=gen) * 10, torch.randn(5, 3, generator=gen) * 10]
    types = [[0, 5, 2], [1, 1, 4, 3, 0]]
    xyz = write_pharmacophore_file(coords, types, cfg["dataset"]["ph_type_map"])
    out = {"hyper_parameters": hp, "config": cfg,
           "xyz_writer": {"coords": [c.tolist() for c in coords], "types": types, "text": xyz},
           "ctor_arg_names": list(hp.keys()),
           "reference": "PharmacophoreDiff.__init__ -> save_hyperparameters (pharmacodiff.py:25-78), configs/dev.yml"}
    path = os.path.join(GOLD, "lightning_ckpt_meta.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(path, os.path.getsize(path), sorted(hp.keys()))


if __name__ == "__main__":
    main()
