"""Shim: LightningModule as a plain nn.Module."""
from torch import nn


class LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


def seed_everything(seed, workers=False):
    import torch

    torch.manual_seed(seed)
