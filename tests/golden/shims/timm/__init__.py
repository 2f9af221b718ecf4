"""Shim of `timm.create_model("tf_efficientnetv2_s_in21ft1k", pretrained=..., features_only=True)` (reference call
site `experiment_modules/bd_model.py:46-51`; timm==0.6.12 per `binarydepth_env.yml:28`, not installed here and not
part of the reference tree).

Test infrastructure only: an independent restatement of the PUBLISHED model definition -- timm's arch strings for
`efficientnetv2_s` decoded below, module names as in timm's `_efficientnet_blocks.py` (`ConvBnAct.conv/bn1`,
`EdgeResidual.conv_exp/bn1/conv_pwl/bn2`, `InvertedResidual.conv_pw/bn1/conv_dw/bn2/se.conv_reduce/se.conv_expand/
conv_pwl/bn3`, `EfficientNetFeatures.conv_stem/bn1/blocks`), TF "SAME" padding (`Conv2dSame`: the `tf_` prefix), BatchNorm
eps 1e-3, SiLU, squeeze width = se_ratio x block input channels.  "Parity unpinned" against the original package
(SURVEY 8c); random-init weights (no checkpoints offline).  The goldens pin the product's encoder
(`implicit_depth_b200/image_encoder.py`) against THIS restatement through the unmodified reference `BDModel`.
"""
import math
import re

import torch
import torch.nn.functional as F
from torch import nn

ARCH_V2_S = [
    ["cn_r2_k3_s1_e1_c24_skip"],
    ["er_r4_k3_s2_e4_c48"],
    ["er_r4_k3_s2_e4_c64"],
    ["ir_r6_k3_s2_e4_c128_se0.25"],
    ["ir_r9_k3_s1_e6_c160_se0.25"],
    ["ir_r15_k3_s2_e6_c256_se0.25"],
]


def _decode(block_str):
    ops = block_str.split("_")
    out = {"type": ops[0], "skip": None, "se": 0.0}
    for op in ops[1:]:
        if op == "skip":
            out["skip"] = True
        elif op == "noskip":
            out["skip"] = False
        else:
            m = re.match(r"([a-z]+)([\d.]+)", op)
            out[m.group(1)] = float(m.group(2)) if "." in m.group(2) else int(m.group(2))
    return out


class _SamePadConv(nn.Conv2d):
    """`Conv2dSame`: pad so that out = ceil(in / stride), extra pixel at the bottom / right."""

    def forward(self, x):
        ih, iw = x.shape[-2:]
        k, s = self.kernel_size[0], self.stride[0]
        ph = max((math.ceil(ih / s) - 1) * s + k - ih, 0)
        pw = max((math.ceil(iw / s) - 1) * s + k - iw, 0)
        if ph > 0 or pw > 0:
            x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
        return F.conv2d(x, self.weight, self.bias, self.stride, 0, self.dilation, self.groups)


def _bn(c):
    return nn.BatchNorm2d(c, eps=1e-3)


class ConvBnAct(nn.Module):
    def __init__(self, cin, cout, k, stride, skip):
        super().__init__()
        self.conv = _SamePadConv(cin, cout, k, stride, bias=False)
        self.bn1 = _bn(cout)
        self.has_skip = bool(skip) and stride == 1 and cin == cout

    def forward(self, x):
        y = F.silu(self.bn1(self.conv(x)))
        return y + x if self.has_skip else y


class EdgeResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, exp, skip):
        super().__init__()
        mid = int(cin * exp)
        self.conv_exp = _SamePadConv(cin, mid, k, stride, bias=False)
        self.bn1 = _bn(mid)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn2 = _bn(cout)
        self.has_skip = skip is not False and stride == 1 and cin == cout

    def forward(self, x):
        y = F.silu(self.bn1(self.conv_exp(x)))
        y = self.bn2(self.conv_pwl(y))
        return y + x if self.has_skip else y


class SqueezeExcite(nn.Module):
    def __init__(self, chs, rd):
        super().__init__()
        self.conv_reduce = nn.Conv2d(chs, rd, 1, bias=True)
        self.conv_expand = nn.Conv2d(rd, chs, 1, bias=True)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(F.silu(self.conv_reduce(s)))
        return x * s.sigmoid()


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, exp, se, skip):
        super().__init__()
        mid = int(cin * exp)
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = _bn(mid)
        self.conv_dw = _SamePadConv(mid, mid, k, stride, groups=mid, bias=False)
        self.bn2 = _bn(mid)
        self.se = SqueezeExcite(mid, int(round(cin * se))) if se > 0 else nn.Identity()
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = _bn(cout)
        self.has_skip = skip is not False and stride == 1 and cin == cout

    def forward(self, x):
        y = F.silu(self.bn1(self.conv_pw(x)))
        y = F.silu(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(self.se(y)))
        return y + x if self.has_skip else y


class _FeatureInfo:
    def __init__(self, chans):
        self._c = list(chans)

    def channels(self):
        return list(self._c)


class EfficientNetFeatures(nn.Module):
    """`features_only=True` wrapper: stem + all stages, returns the last output of each resolution."""

    def __init__(self, arch, stem_size=24):
        super().__init__()
        self.conv_stem = _SamePadConv(3, stem_size, 3, 2, bias=False)
        self.bn1 = _bn(stem_size)
        stages, cin, reduction = [], stem_size, 2
        self._taps, chans = [], []
        for si, stage in enumerate(arch):
            blocks = []
            for bstr in stage:
                a = _decode(bstr)
                for r in range(a["r"]):
                    stride = a["s"] if r == 0 else 1
                    if a["type"] == "cn":
                        blocks.append(ConvBnAct(cin, a["c"], a["k"], stride, a["skip"]))
                    elif a["type"] == "er":
                        blocks.append(EdgeResidual(cin, a["c"], a["k"], stride, a["e"], a["skip"]))
                    elif a["type"] == "ir":
                        blocks.append(InvertedResidual(cin, a["c"], a["k"], stride, a["e"], a["se"], a["skip"]))
                    else:
                        raise NotImplementedError(a["type"])
                    cin = a["c"]
                reduction *= a["s"]
            stages.append(nn.Sequential(*blocks))
            # a feature is taken where the NEXT stage reduces the resolution (and at the end)
            nxt = arch[si + 1] if si + 1 < len(arch) else None
            if nxt is None or _decode(nxt[0])["s"] > 1:
                self._taps.append(si)
                chans.append(cin)
        self.blocks = nn.Sequential(*stages)
        self.feature_info = _FeatureInfo(chans)

    def forward(self, x):
        x = F.silu(self.bn1(self.conv_stem(x)))
        outs = []
        for i, stage in enumerate(self.blocks):
            x = stage(x)
            if i in self._taps:
                outs.append(x)
        return outs


def create_model(name, pretrained=False, features_only=True, **kw):
    if "efficientnetv2_s" in name:
        assert name.startswith("tf_"), "only the TF-padding variant the reference uses is restated"
        return EfficientNetFeatures(ARCH_V2_S)
    raise NotImplementedError(f"timm shim: {name}")
