"""Checkpoint and run-directory I/O in the reference's formats, without pytorch_lightning.

  * `load_run`: what generate_pharmacophores.py:231-269 does before sampling -- locate `config.yaml` / `config.yml` and
    `checkpoints/last.ckpt` from a run directory or a checkpoint path, parse the config, build the model from the
    checkpoint's `hyper_parameters`, and retry with `ph_type_map` from the config when the checkpoint predates that
    constructor argument (the reference's `except TypeError` branch, :264-268).
  * `write_run_dir`: the run directory train.py:114-130 creates -- `<output_dir>/<name>_<run_id>/config.yaml` carrying
    `resume.run_id` and `wandb.name`, plus `checkpoints/`.
  * `save_lightning_checkpoint`: a `.ckpt` with the keys Lightning's ModelCheckpoint writes (`state_dict`,
    `hyper_parameters`, `epoch`, `global_step`, `pytorch-lightning_version`), so the reference can load what this
    package trains and vice versa.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple

import torch
import yaml

from .diffusion import PharmacophoreDiff


def find_run_files(model_dir: Optional[Path] = None, ckpt: Optional[Path] = None) -> Tuple[Path, Path]:
    """-> (config_file, model_file), generate_pharmacophores.py:231-246."""
    if ckpt is not None:
        ckpt = Path(ckpt)
        run_dir, model_file = ckpt.parent.parent, ckpt
    elif model_dir is not None:
        run_dir = Path(model_dir)
        model_file = run_dir / "checkpoints" / "last.ckpt"
    else:
        raise ValueError("either model_dir or ckpt must be given")
    config_file = run_dir / "config.yaml"
    if not config_file.exists():
        config_file = run_dir / "config.yml"
        if not config_file.exists():
            raise FileNotFoundError(f"config file not found in {run_dir}")
    return config_file, model_file


def load_run(model_dir: Optional[Path] = None, ckpt: Optional[Path] = None, device=None):
    """-> (model in eval mode, config dict).  The TypeError retry mirrors generate_pharmacophores.py:264-268."""
    config_file, model_file = find_run_files(model_dir, ckpt)
    with open(config_file) as f:
        config = yaml.load(f, Loader=yaml.FullLoader)
    try:
        model = PharmacophoreDiff.load_from_checkpoint(model_file)
    except TypeError:
        model = PharmacophoreDiff.load_from_checkpoint(model_file, ph_type_map=config["dataset"]["ph_type_map"])
    if device is not None:
        model = model.to(device)
    model.eval()
    return model, config


def write_run_dir(output_dir, config: dict, name: str, run_id: str) -> Path:
    """train.py:114-130: the run directory with the resumable config; returns it."""
    config = dict(config)
    config["resume"] = {"run_id": run_id}
    config["wandb"] = dict(config.get("wandb") or {}, name=name)
    run_dir = Path(output_dir) / f"{name}_{run_id}"
    (run_dir / "checkpoints").mkdir(parents=True, exist_ok=True)
    with open(run_dir / "config.yaml", "w") as f:
        yaml.dump(config, f)
    return run_dir


def save_lightning_checkpoint(model: PharmacophoreDiff, path, epoch: int = 0, global_step: int = 0,
                              optimizer: Optional[torch.optim.Optimizer] = None):
    ckpt = {"epoch": epoch, "global_step": global_step, "pytorch-lightning_version": "2.0.0",
            "state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "hyper_parameters": dict(model.hparams)}
    if optimizer is not None:
        ckpt["optimizer_states"] = [optimizer.state_dict()]
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    torch.save(ckpt, path)
