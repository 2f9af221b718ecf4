"""Shim of the four kornia.filters functions the reference names.

Real source files are required because the reference TorchScript-compiles a
caller (`utils/generic_utils.py:84-91`)."""
from typing import Tuple

import torch
import torch.nn.functional as F


def blur_pool2d(input: torch.Tensor, kernel_size: int, stride: int = 2) -> torch.Tensor:
    k = torch.tensor([1.0, 2.0, 1.0], dtype=input.dtype, device=input.device)
    k2 = (k[:, None] * k[None, :]) / 16.0
    c = input.shape[1]
    w = k2[None, None].repeat(c, 1, 1, 1)
    return F.conv2d(input, w, stride=stride, padding=1, groups=c)


def gaussian_blur2d(input: torch.Tensor, kernel_size: Tuple[int, int], sigma: Tuple[float, float]) -> torch.Tensor:
    ks = kernel_size[0]
    x = torch.arange(ks, dtype=input.dtype, device=input.device) - (ks - 1) / 2.0
    g = torch.exp(-(x * x) / (2.0 * sigma[0] * sigma[0]))
    g = g / g.sum()
    k2 = g[:, None] * g[None, :]
    c = input.shape[1]
    w = k2[None, None].repeat(c, 1, 1, 1)
    pad = ks // 2
    return F.conv2d(F.pad(input, [pad, pad, pad, pad], mode="reflect"), w, groups=c)


def spatial_gradient(input: torch.Tensor) -> torch.Tensor:
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=input.dtype, device=input.device) / 8.0
    ky = kx.t()
    b, c, h, w = input.shape
    x = F.pad(input.reshape(b * c, 1, h, w), [1, 1, 1, 1], mode="replicate")
    gx = F.conv2d(x, kx[None, None])
    gy = F.conv2d(x, ky[None, None])
    return torch.stack([gx, gy], dim=2).reshape(b, c, 2, h, w)


def sobel(input: torch.Tensor) -> torch.Tensor:
    g = spatial_gradient(input)
    return torch.sqrt(g[:, :, 0] ** 2 + g[:, :, 1] ** 2 + 1e-6)
