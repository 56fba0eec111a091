"""Minimal pytorch_lightning surface for importing pharmacodiff.py unmodified.

TEST INFRASTRUCTURE ONLY.  LightningModule here is an nn.Module with
save_hyperparameters()/log_dict()/device; no Trainer, no checkpoint IO.
"""
import inspect

import torch
import torch.nn as nn


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.hparams = {}
        self.trainer = None
        self.current_epoch = 0
        self.logged = []

    def save_hyperparameters(self, *args, **kwargs):
        frame = inspect.currentframe().f_back
        init_args = {}
        local_vars = frame.f_locals
        sig = inspect.signature(type(self).__init__)
        for name, p in sig.parameters.items():
            if name == "self":
                continue
            if p.kind == inspect.Parameter.VAR_KEYWORD:
                init_args.update(local_vars.get(name, {}))
            elif name in local_vars:
                init_args[name] = local_vars[name]
        self.hparams = init_args

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def log_dict(self, d, **kwargs):
        self.logged.append(dict(d))

    def log(self, k, v, **kwargs):
        self.logged.append({k: v})


class LightningDataModule:
    pass


def seed_everything(seed, workers=False):
    torch.manual_seed(seed)
    return seed
