"""Output side of the reference's evaluation scripts on one kernel.

`sigmoid_upsample(pred_0, size, multiplier, mode)` = `F.interpolate(sigmoid_custom(pred_0, multiplier), size=size,
mode=mode)` of `test_bd.py:225-243` / `inference/inference.py:159-162`; `resize(x, size, mode)` is the plain
`F.interpolate` the scripts apply to `rendered_depth` / `search_depths` (`test_bd.py:245-271`)."""
from __future__ import annotations

import torch

from . import _abi


def _run(x, size, multiplier, mode, apply_sigmoid):
    if mode not in ("bilinear", "nearest"):
        raise ValueError(f"mode must be 'bilinear' or 'nearest', got {mode!r}")
    _abi.require_cuda(x)
    if x.dim() != 4:
        raise ValueError(f"expected a B x P x h x w tensor, got {tuple(x.shape)}")
    x = (x if x.dtype == torch.float32 else x.float()).contiguous()
    B, P, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((B, P, H, W), device=x.device, dtype=torch.float32)
    _abi.call("b200_sigmoid_resize", _abi.ptr(x), _abi.ptr(out), B * P, h, w, H, W, float(multiplier),
              1 if mode == "nearest" else 0, 1 if apply_sigmoid else 0, _abi.stream_ptr())
    return out


def sigmoid_upsample(pred, size, multiplier=1.0, mode="bilinear"):
    return _run(pred, size, multiplier, mode, True)


def resize(x, size, mode="nearest"):
    return _run(x, size, 1.0, mode, False)
