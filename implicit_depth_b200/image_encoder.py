"""EfficientNetV2-S image-prior encoder on the hand-written sm_100a kernels.

The reference takes its image encoder from timm (`tf_efficientnetv2_s_in21ft1k`, features_only, bd_model.py:46-51; a
third-party backbone).  cuDNN runs it either in TF32 -- which alone moves `pred_0` 1.7e-2
away from the fp32 reference, outside the 1e-3 parity budget -- or in strict fp32 at 9.5 ms per batch of four
frames, longer than the whole rest of the forward.  This module runs the same layers (torchvision `efficientnet_v2_s`
layout, the stand-in used throughout this repo) on the split-bf16 tensor-core conv kernels (fp32-grade results) plus
three small MBConv kernels (csrc/mbconv.cu): depthwise 3x3, squeeze-excite, scale.

Channel counts that the conv kernels cannot produce (24, 160, 192, 960: Cout must be a multiple of 16 up to 128 and
of 128 beyond) are zero-padded; padded channels stay exactly zero through SiLU / depthwise / SE / residuals and the
consumers' weights are padded with zero columns (`SplitAct.Cl` = logical channels).
"""
from __future__ import annotations


import torch

from . import _abi
from .networks import Plan, _fold_bn

FUSED_DW_SE = True  # False = separate depthwise / pool / fc kernels (kept for the kernel-level tests)


def pad_ch(c):
    return (c + 15) // 16 * 16 if c <= 128 else (c + 127) // 128 * 128


def _padw(w, cout_p, cin_p):
    """[Cout, Cin, k, k] -> zero-padded [cout_p, cin_p, k, k] (fp32)."""
    out = torch.zeros((cout_p, cin_p) + tuple(w.shape[2:]), dtype=torch.float32, device=w.device)
    out[:w.shape[0], :w.shape[1]] = w.detach().float()
    return out


def _padv(v, n):
    out = torch.zeros(n, dtype=torch.float32, device=v.device)
    out[:v.numel()] = v.detach().float().reshape(-1)
    return out


def _cna(g: Plan, x, cna, act, stride=1, residual=None):
    """torchvision Conv2dNormActivation (conv, BatchNorm[, SiLU]) with BN folded, on the conv kernels."""
    conv, bn = cna[0], cna[1]
    w, b = _fold_bn(conv.weight, bn)
    cout = conv.out_channels
    cp = pad_ch(cout)
    k = conv.kernel_size[0]
    y, _ = g.conv([(x, _padw(w.to(g.device), cp, x.C), stride, k // 2)], _padv(b.to(g.device), cp), cp, act=act,
                  residual=residual)
    y.Cl = cout
    return y


def _dwconv(g: Plan, x, cna, stride):
    conv, bn = cna[0], cna[1]
    w, b = _fold_bn(conv.weight, bn)  # [C, 1, 3, 3]
    C = x.C
    wt = torch.zeros((9, C), dtype=torch.float32, device=g.device)
    wt[:, :w.shape[0]] = w.to(g.device).float().reshape(w.shape[0], 9).t()
    bias = _padv(b.to(g.device), C)
    OH, OW = (x.H + 2 - 3) // stride + 1, (x.W + 2 - 3) // stride + 1
    y = g.act(x.B, OH, OW, C)
    y.Cl = getattr(x, "Cl", C)
    g._keep += [wt, bias]
    g.add(lambda: _abi.call("b200_dwconv3x3_silu", _abi.ptr(x.hi), _abi.ptr(x.lo), _abi.ptr(wt), _abi.ptr(bias),
                            _abi.ptr(y.hi), _abi.ptr(y.lo), x.B, x.H, x.W, C, stride, _abi.stream_ptr()),
          reads=[x], writes=[y])
    return y


def _squeeze_excite(g: Plan, x, se):
    """torchvision SqueezeExcitation: avgpool -> fc1 -> SiLU -> fc2 -> sigmoid -> scale (in place)."""
    C, S = x.C, se.fc1.out_channels
    cl = se.fc1.in_channels
    w1 = torch.zeros((S, C), dtype=torch.float32, device=g.device)
    w1[:, :cl] = se.fc1.weight.detach().to(g.device).float().reshape(S, cl)
    b1 = se.fc1.bias.detach().to(g.device).float().contiguous()
    w2 = torch.zeros((S, C), dtype=torch.float32, device=g.device)  # transposed: coalesced over channels
    w2[:, :cl] = se.fc2.weight.detach().to(g.device).float().reshape(cl, S).t()
    b2 = _padv(se.fc2.bias.to(g.device), C)
    mean = g.empty((x.B, C))
    scale = g.empty((x.B, C))
    g._keep += [w1, b1, w2, b2]
    g.add(lambda: _abi.call("b200_squeeze_excite", _abi.ptr(x.hi), _abi.ptr(x.lo), _abi.ptr(w1), _abi.ptr(b1),
                            _abi.ptr(w2), _abi.ptr(b2), _abi.ptr(mean), _abi.ptr(scale), _abi.ptr(x.hi),
                            _abi.ptr(x.lo), x.B, x.H * x.W, C, S, _abi.stream_ptr()), launches=3, reads=[x], writes=[x])
    return x


def _dwconv_se(g: Plan, x, cna, stride, se):
    """MBConv middle: depthwise 3x3 + BN + SiLU, then SqueezeExcitation, as one chain of four short kernels with the
    squeeze fused into the depthwise conv (`b200_mbconv_dw_se`)."""
    conv, bn = cna[0], cna[1]
    w, b = _fold_bn(conv.weight, bn)  # [C, 1, 3, 3]
    C, S, cl = x.C, se.fc1.out_channels, se.fc1.in_channels
    wt = torch.zeros((9, C), dtype=torch.float32, device=g.device)
    wt[:, :w.shape[0]] = w.to(g.device).float().reshape(w.shape[0], 9).t()
    bias = _padv(b.to(g.device), C)
    w1 = torch.zeros((S, C), dtype=torch.float32, device=g.device)
    w1[:, :cl] = se.fc1.weight.detach().to(g.device).float().reshape(S, cl)
    b1 = se.fc1.bias.detach().to(g.device).float().contiguous()
    w2 = torch.zeros((S, C), dtype=torch.float32, device=g.device)  # transposed: coalesced over channels
    w2[:, :cl] = se.fc2.weight.detach().to(g.device).float().reshape(cl, S).t()
    b2 = _padv(se.fc2.bias.to(g.device), C)
    OH, OW = (x.H + 2 - 3) // stride + 1, (x.W + 2 - 3) // stride + 1
    pix = _abi.load().b200_mbconv_pool_block()
    partial = g.empty((x.B, (OH * OW + pix - 1) // pix, C))
    s1 = g.empty((x.B, S))
    scale = g.empty((x.B, C))
    y = g.act(x.B, OH, OW, C)
    y.Cl = getattr(x, "Cl", C)
    g._keep += [wt, bias, w1, b1, w2, b2]
    g.add(lambda: _abi.call("b200_mbconv_dw_se", _abi.ptr(x.hi), _abi.ptr(x.lo), _abi.ptr(wt), _abi.ptr(bias),
                            _abi.ptr(w1), _abi.ptr(b1), _abi.ptr(w2), _abi.ptr(b2), _abi.ptr(partial), _abi.ptr(s1),
                            _abi.ptr(scale), _abi.ptr(y.hi), _abi.ptr(y.lo), x.B, x.H, x.W, C, stride, S,
                            _abi.stream_ptr()), launches=4, reads=[x], writes=[y])
    return y


def plan_efficientnet_v2_s(g: Plan, features, get_image, B, H, W, taps=(1, 2, 3, 5, 6)):
    """Launch plan of torchvision `efficientnet_v2_s().features[:7]` (stem + 6 stages).  Returns the SplitActs of the
    tapped stages (channels [24, 48, 64, 160, 256] at /2 .. /32; `.Cl` holds the logical channel count)."""
    from torchvision.models.efficientnet import FusedMBConv, MBConv

    dev = g.device
    img8 = torch.zeros((B, 8, H, W), device=dev, dtype=torch.float32)  # channels 3..7 stay zero
    g._keep.append(img8)

    def load():
        img = get_image()
        assert tuple(img.shape) == (B, 3, H, W), f"expected {(B, 3, H, W)}, got {tuple(img.shape)}"
        img8[:, :3].copy_(img)
        return img8

    x = g.from_f32(load, B, 8, H, W)
    x = _cna(g, x, features[0], "silu", stride=2)
    outs = []
    for si in range(1, len(features)):
        for blk in features[si]:
            inp = x
            res = inp if blk.use_res_connect else None
            if isinstance(blk, FusedMBConv):
                layers = blk.block
                stride = layers[0][0].stride[0]
                if len(layers) == 1:  # expand ratio 1: a single 3x3 conv + SiLU
                    x = _cna(g, inp, layers[0], "silu", stride, residual=None)
                    if res is not None:  # the residual is added AFTER the activation: not the conv epilogue's order
                        x = _add(g, x, res)
                else:
                    h = _cna(g, inp, layers[0], "silu", stride)
                    x = _cna(g, h, layers[1], "none", 1, residual=res)
            elif isinstance(blk, MBConv):
                layers = blk.block
                stride = layers[1][0].stride[0]
                h = _cna(g, inp, layers[0], "silu", 1)
                if FUSED_DW_SE:
                    h = _dwconv_se(g, h, layers[1], stride, layers[2])
                else:
                    h = _dwconv(g, h, layers[1], stride)
                    h = _squeeze_excite(g, h, layers[2])
                x = _cna(g, h, layers[3], "none", 1, residual=res)
            else:
                raise TypeError(f"unsupported block {type(blk).__name__}")
        if si in taps:
            outs.append(x)
    return outs


def _add(g: Plan, a, b):
    """a + b on split activations (FusedMBConv with expand ratio 1 adds its input after the SiLU)."""
    out = g.act(a.B, a.H, a.W, a.C)
    out.Cl = getattr(a, "Cl", a.C)
    n = a.B * a.H * a.W * a.C
    g.add(lambda: _abi.call("b200_split_add", _abi.ptr(a.hi), _abi.ptr(a.lo), _abi.ptr(b.hi), _abi.ptr(b.lo),
                            _abi.ptr(out.hi), _abi.ptr(out.lo), n, _abi.stream_ptr()), reads=[a, b], writes=[out])
    return out
