"""MBConv middle of the image-prior encoder (`b200_mbconv_dw_se`: depthwise 3x3 + SiLU with the squeeze fused in,
fc1, fc2, scale) against torch fp32 on the same inputs, including ragged sizes (pixels not a multiple of the pool
block, channels not a multiple of 64, stride 2), and against the separate-kernel chain it replaces."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from implicit_depth_b200 import _abi
from implicit_depth_b200.conv import SplitAct

from cases import rel_err

pytestmark = pytest.mark.gpu


def _inputs(seed, B, C, H, W, S):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return dict(x=r(B, C, H, W), wd=r(C, 1, 3, 3) / 3, bd=0.1 * r(C), w1=r(S, C) / C ** 0.5, b1=0.1 * r(S),
                w2=r(C, S) / S ** 0.5, b2=0.1 * r(C))


def _torch_ref(t, stride, pad_lo=1):
    x = SplitAct.from_nchw_torch(t["x"]).float_nchw()  # what the kernels actually read (hi + lo)
    C = x.shape[1]
    xp = F.pad(x.double(), [pad_lo, 1, pad_lo, 1])  # (1, 1): torch pad 1; (0, 1): TF "SAME" at stride 2, even sizes
    y = F.silu(F.conv2d(xp, t["wd"].double(), t["bd"].double(), stride=stride, padding=0, groups=C))
    m = y.mean((2, 3))
    s1 = F.silu(m @ t["w1"].double().t() + t["b1"].double())
    sc = torch.sigmoid(s1 @ t["w2"].double().t() + t["b2"].double())
    return (y * sc[:, :, None, None]).float()


def _run_fused(t, stride, pad_lo=1):
    B, C, H, W = t["x"].shape
    S = t["w1"].shape[0]
    c = lambda v: v.cuda().float().contiguous()
    x = SplitAct.from_nchw_torch(t["x"].cuda())
    wt = c(t["wd"].reshape(C, 9).t())
    bias, w1, b1, w2t, b2 = c(t["bd"]), c(t["w1"]), c(t["b1"]), c(t["w2"].t()), c(t["b2"])
    OH, OW = (H + pad_lo + 1 - 3) // stride + 1, (W + pad_lo + 1 - 3) // stride + 1
    pix = _abi.load().b200_mbconv_pool_block()
    partial = torch.full((B, (OH * OW + pix - 1) // pix, C), float("nan"), device="cuda")
    s1 = torch.full((B, S), float("nan"), device="cuda")
    scale = torch.full((B, C), float("nan"), device="cuda")
    y = SplitAct(B, OH, OW, C, "cuda")
    _abi.call("b200_mbconv_dw_se", _abi.ptr(x.hi), _abi.ptr(x.lo), _abi.ptr(wt), _abi.ptr(bias), _abi.ptr(w1),
              _abi.ptr(b1), _abi.ptr(w2t), _abi.ptr(b2), _abi.ptr(partial), _abi.ptr(s1), _abi.ptr(scale),
              _abi.ptr(y.hi), _abi.ptr(y.lo), B, H, W, C, stride, S, pad_lo, _abi.stream_ptr())
    torch.cuda.synchronize()
    return y.float_nchw().cpu(), (x, wt, bias, w1, b1, w2t, b2)


@pytest.mark.parametrize("shape", [(2, 64, 12, 16, 1, 16), (1, 72, 7, 9, 1, 5), (3, 136, 13, 11, 2, 34),
                                   (4, 1536, 12, 16, 1, 64), (2, 960, 48, 64, 2, 40), (1, 8, 1, 1, 1, 1)])
def test_mbconv_dw_se_vs_torch(shape):
    B, C, H, W, stride, S = shape
    t = _inputs(100 + C, B, C, H, W, S)
    got, _ = _run_fused(t, stride)
    ref = _torch_ref(t, stride)
    assert got.shape == ref.shape
    assert rel_err(got.numpy(), ref.numpy()) < 2e-5  # fp32 arithmetic + one split-bf16 rounding (2^-17)


@pytest.mark.parametrize("shape", [(2, 256, 24, 32, 16), (1, 768, 12, 16, 32), (3, 72, 6, 10, 5)])
def test_mbconv_dw_se_tf_same_padding_stride2(shape):
    """TF "SAME" padding of the stride-2 depthwise convs of timm's tf_efficientnetv2_s (stages 3 and 5) on even-sized
    maps: no padding at the top / left, one pixel at the bottom / right."""
    B, C, H, W, S = shape
    t = _inputs(300 + C, B, C, H, W, S)
    got, _ = _run_fused(t, 2, pad_lo=0)
    ref = _torch_ref(t, 2, pad_lo=0)
    assert got.shape == ref.shape == (B, C, H // 2, W // 2)
    assert rel_err(got.numpy(), ref.numpy()) < 2e-5


def test_mbconv_dw_se_matches_separate_kernels_and_is_batch_invariant():
    B, C, H, W, stride, S = 3, 512, 24, 32, 1, 32
    t = _inputs(7, B, C, H, W, S)
    got, (x, wt, bias, w1, b1, w2t, b2) = _run_fused(t, stride)
    y = SplitAct(B, H, W, C, "cuda")
    _abi.call("b200_dwconv3x3_silu", _abi.ptr(x.hi), _abi.ptr(x.lo), _abi.ptr(wt), _abi.ptr(bias), _abi.ptr(y.hi),
              _abi.ptr(y.lo), B, H, W, C, stride, 1, _abi.stream_ptr())
    mean, scale = torch.empty((B, C), device="cuda"), torch.empty((B, C), device="cuda")
    _abi.call("b200_squeeze_excite", _abi.ptr(y.hi), _abi.ptr(y.lo), _abi.ptr(w1), _abi.ptr(b1), _abi.ptr(w2t),
              _abi.ptr(b2), _abi.ptr(mean), _abi.ptr(scale), _abi.ptr(y.hi), _abi.ptr(y.lo), B, H * W, C, S,
              _abi.stream_ptr())
    torch.cuda.synchronize()
    assert rel_err(got.numpy(), y.float_nchw().cpu().numpy()) < 2e-5
    # frame 1 alone gives bit-identical results (fixed-order reductions, no cross-frame state)
    t1 = {k: (v[1:2] if k == "x" else v) for k, v in t.items()}
    alone, _ = _run_fused(t1, stride)
    assert torch.equal(alone[0], got[1])


def test_mbconv_dw_se_bad_arguments():
    lib = _abi.load()
    assert lib.b200_mbconv_dw_se(*([None] * 13), 1, 4, 4, 8, 1, 4, 1, None) == -1
    d = torch.zeros(64, device="cuda")
    p = _abi.ptr(d)
    assert lib.b200_mbconv_dw_se(*([p] * 13), 1, 4, 4, 12, 1, 4, 1, None) == -1   # C % 8
    assert lib.b200_mbconv_dw_se(*([p] * 13), 1, 4, 4, 8, 3, 4, 1, None) == -1    # stride
    assert lib.b200_mbconv_dw_se(*([p] * 13), 1, 4, 4, 8, 1, 200, 1, None) == -1  # S > 128
    assert lib.b200_mbconv_dw_se(*([p] * 13), 1, 4, 4, 8, 2, 4, 2, None) == -1    # pad_lo


@pytest.mark.parametrize("pad_lo", [0, 1])
@pytest.mark.parametrize("shape", [(2, 24, 32, 24), (1, 37, 45, 32), (3, 6, 10, 8)])
def test_encoder_stem_direct_conv_vs_torch(shape, pad_lo):
    """`b200_stem3x3_s2_silu` (image-encoder stem: Conv2d(3, Cout, 3, stride 2) + bias + SiLU straight from the fp32
    image) vs torch fp64: TF "SAME" padding (0, 1) and symmetric padding 1, ragged sizes, channel padding, and an
    image that is a strided view (the forward passes slices of a staging buffer)."""
    B, H, W, cout = shape
    cp = (cout + 15) // 16 * 16
    g = torch.Generator().manual_seed(B * 100 + H + pad_lo)
    big = torch.randn(B, 4, H, W + 3, generator=g).cuda()
    img = big[:, :3, :, 1:W + 1]                      # non-contiguous view
    w = (torch.randn(cout, 3, 3, 3, generator=g) / 27 ** 0.5).cuda()
    b = (0.1 * torch.randn(cout, generator=g)).cuda()
    wk = torch.zeros(27, cp, device="cuda")
    wk[:, :cout] = w.reshape(cout, 27).t()
    bias = torch.zeros(cp, device="cuda")
    bias[:cout] = b
    OH, OW = (H + pad_lo + 1 - 3) // 2 + 1, (W + pad_lo + 1 - 3) // 2 + 1
    out = SplitAct(B, OH, OW, cp, "cuda")
    out.hi.fill_(float("nan")); out.lo.fill_(float("nan"))
    _abi.call("b200_stem3x3_s2_silu", _abi.ptr(img), _abi.ptr(wk), _abi.ptr(bias), _abi.ptr(out.hi), _abi.ptr(out.lo),
              B, H, W, cp, pad_lo, img.stride(0), img.stride(1), img.stride(2), img.stride(3), _abi.stream_ptr())
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(F.pad(img.double(), [pad_lo, 1, pad_lo, 1]), w.double(), b.double(), stride=2))
    got = out.float_nchw().double()
    assert torch.isfinite(got).all()
    assert got.shape == (B, cp, OH, OW)
    assert (got[:, :cout] - ref).abs().max().item() < 3e-5 * ref.abs().max().item()
    assert (got[:, cout:] == 0).all()  # padded channels: zero weights and bias -> silu(0) = 0
