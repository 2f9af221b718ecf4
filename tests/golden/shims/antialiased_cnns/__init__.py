"""Shim of antialiased-cnns==0.3 `resnet18` (published algorithm: Zhang, "Making
Convolutional Networks Shift-Invariant Again", ICML 2019; package resnet.py/blurpool.py).
Only the stem the reference uses (`modules/networks.py:250-270`): conv1, bn1, relu,
maxpool (= MaxPool2d(2, stride 1) + BlurPool(filt 4, stride 2)), layer1."""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class BlurPool(nn.Module):
    def __init__(self, channels, filt_size=4, stride=2):
        super().__init__()
        self.pad_sizes = [int(1.0 * (filt_size - 1) / 2), int(np.ceil(1.0 * (filt_size - 1) / 2))] * 2
        self.stride = stride
        self.channels = channels
        a = {1: [1.0], 2: [1.0, 1.0], 3: [1.0, 2.0, 1.0], 4: [1.0, 3.0, 3.0, 1.0], 5: [1.0, 4.0, 6.0, 4.0, 1.0]}[filt_size]
        a = torch.tensor(a)
        filt = a[:, None] * a[None, :]
        filt = filt / filt.sum()
        self.register_buffer("filt", filt[None, None].repeat(channels, 1, 1, 1))
        self.pad = nn.ReflectionPad2d(self.pad_sizes)

    def forward(self, x):
        return F.conv2d(self.pad(x), self.filt, stride=self.stride, groups=x.shape[1])


class BasicBlock(nn.Module):
    def __init__(self, inplanes, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + x)


class _ResNetStem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=1), BlurPool(64, filt_size=4, stride=2))
        self.layer1 = nn.Sequential(BasicBlock(64, 64), BasicBlock(64, 64))


def resnet18(pretrained=False, **kw):
    return _ResNetStem()


resnet34 = resnet50 = resnet101 = resnet152 = resnet18
