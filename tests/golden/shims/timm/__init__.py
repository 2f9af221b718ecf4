"""Shim: `timm.create_model(..., features_only=True)` -> torchvision EfficientNetV2-S
feature taps with the same channel/stride layout ([24,48,64,160,256] at /2../32)
as `tf_efficientnetv2_s_in21ft1k` (reference call site `experiment_modules/bd_model.py:46-51`).
Random-init weights (no checkpoints offline)."""
import torch
from torch import nn


class _FeatureInfo:
    def __init__(self, chans):
        self._c = list(chans)

    def channels(self):
        return list(self._c)


class EffNetV2SFeatures(nn.Module):
    TAPS = (1, 2, 3, 5, 6)

    def __init__(self):
        super().__init__()
        import torchvision

        net = torchvision.models.efficientnet_v2_s(weights=None)
        self.features = nn.Sequential(*list(net.features)[:7])
        self.feature_info = _FeatureInfo([24, 48, 64, 160, 256])

    def forward(self, x):
        outs = []
        for i, m in enumerate(self.features):
            x = m(x)
            if i in self.TAPS:
                outs.append(x)
        return outs


def create_model(name, pretrained=False, features_only=True, **kw):
    if "efficientnetv2_s" in name:
        return EffNetV2SFeatures()
    raise NotImplementedError(f"timm shim: {name}")
