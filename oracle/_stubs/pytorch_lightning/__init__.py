"""Minimal stand-in for pytorch_lightning, ONLY used by oracle/ref_loader.py to import the
reference's hand-written ViT classes unmodified in a container that lacks Lightning.
(vision_transformer_base.py:11 imports it for the LightningModule base class only.)"""
import inspect

import torch.nn as nn


class _HP(dict):
    __getattr__ = dict.get


class LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        loc = inspect.currentframe().f_back.f_locals
        hp = {n: v for n, v in loc.items() if n not in ("self", "__class__", "kwargs")}
        hp.update(loc.get("kwargs", {}))
        object.__setattr__(self, "_hp", _HP(hp))

    @property
    def hparams(self):
        return self._hp

    def log(self, *a, **k):
        pass


def seed_everything(s, workers=False):
    import random

    import numpy as np
    import torch
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)
    return s
