"""Tensor-core convolution kernel vs torch fp64 conv2d (test-side reference of the same op)."""
import pytest
import torch
import torch.nn.functional as F

from implicit_depth_b200.conv import ConvPlan, SplitAct

pytestmark = pytest.mark.gpu


def act_ref(x, act, slope):
    if act == "lrelu":
        return F.leaky_relu(x, slope)
    if act == "elu":
        return F.elu(x)
    if act == "relu":
        return F.relu(x)
    if act == "silu":
        return F.silu(x)
    return x


# (B, H, W, [segment channel counts], Cout, ksize, stride, act, residual)
CASES = [
    (1, 8, 16, [64], 64, 3, 1, "none", False),
    (2, 24, 32, [64], 64, 3, 1, "lrelu", True),
    (1, 16, 16, [24], 64, 3, 1, "lrelu", False),      # channel tail via TMA OOB fill
    (1, 12, 16, [64, 48], 64, 3, 1, "lrelu", False),  # concat as two K-segments
    (2, 15, 20, [128], 128, 3, 1, "elu", False),      # ragged spatial size
    (1, 16, 32, [64], 128, 1, 1, "none", False),      # 1x1
    (1, 16, 32, [64], 128, 3, 2, "lrelu", False),     # stride 2 through tensor-map element strides
    (1, 24, 32, [160, 256], 256, 3, 1, "lrelu", False),  # two N tiles
    (1, 8, 16, [128], 16, 3, 1, "none", False),       # narrow N (matching head)
    (3, 6, 8, [384, 256], 384, 3, 1, "relu", False),  # coarsest level, three N tiles
    (2, 37, 45, [128], 16, 3, 1, "lrelu", False),     # narrow N, ragged size, several items per CTA
    (5, 96, 128, [128], 16, 3, 1, "none", False),     # narrow N, M=256 double tiles on a full grid
    (2, 24, 32, [256], 256, 3, 1, "lrelu", True),     # cluster split-K x4 with a residual (32 items, 8 patches)
    (1, 13, 19, [192], 128, 3, 1, "elu", True),       # split-K x2, ragged map: border rows / columns of the slice stores
    (4, 12, 16, [384, 256], 384, 3, 1, "lrelu", False),  # CVEncoder level 3 at cfg2: 24 items x 4 CTAs
    (1, 24, 32, [512], 64, 3, 1, "relu", True),       # 64-wide N tile, split-K x8 (8 items, 16 patches), residual
    (2, 37, 45, [24], 32, 3, 1, "silu", False),       # 32-wide narrow N tile (image-encoder stage 1), ragged size
    (4, 96, 128, [32], 32, 3, 1, "silu", False),      # 32-wide narrow N tile, M=256 double tiles on a full grid
    (2, 13, 19, [96], 128, 1, 1, "elu", True),        # plain kernel, staged (transposed) stores: ragged map, residual
    (1, 11, 21, [40], 64, 1, 1, "relu", False),       # plain kernel, staged stores with the 64-wide N tile
    (2, 21, 27, [64], 256, 3, 2, "lrelu", False),     # stride 2, two staged N tiles, ragged output (11 x 14)
]


@pytest.mark.parametrize("with_f32", [False, True])  # an fp32 copy forces the per-tap kernel, else halo if eligible
@pytest.mark.parametrize("case", CASES)
def test_conv_matches_fp64(case, with_f32):
    B, H, W, segC, Cout, k, stride, act, use_res = case
    torch.manual_seed(sum(segC) + Cout + k + stride)
    xs = [torch.randn(B, C, H, W, device="cuda") for C in segC]
    ws = [torch.randn(Cout, C, k, k, device="cuda") / (C * k * k) ** 0.5 for C in segC]
    bias = torch.randn(Cout, device="cuda")
    pad = k // 2
    acts = [SplitAct.from_nchw_torch(x) for x in xs]
    OH = (H + 2 * pad - k) // stride + 1
    OW = (W + 2 * pad - k) // stride + 1
    out = SplitAct(B, OH, OW, Cout, "cuda")
    out.hi.fill_(float("nan")); out.lo.fill_(float("nan"))
    out_f32 = torch.full((B, OH, OW, Cout), float("nan"), device="cuda") if with_f32 else None
    res = None
    if use_res:
        rx = torch.randn(B, Cout, OH, OW, device="cuda")
        res = SplitAct.from_nchw_torch(rx)
    plan = ConvPlan([(a, k, stride, pad) for a in acts], ws, bias, out, B, Cout, act=act, slope=0.2,
                    residual=res, out_f32=out_f32)
    # pure 1x1 convs and fp32 copies: per-tap kernel; Cout == 16 / 32 without residual: the narrow halo variants
    assert plan.halo == (not with_f32 and stride == 1 and k == 3 and (Cout % 64 == 0 or (Cout in (16, 32) and not use_res)))
    plan.run()
    plan.run()  # a second launch must give the same answer (persistent state fully re-initialised)
    torch.cuda.synchronize()
    ref = sum(F.conv2d(a.float_nchw().double(), w.double(), None, stride, pad) for a, w in zip(acts, ws))
    ref = ref + bias.double().view(1, -1, 1, 1)
    if use_res:
        ref = ref + res.float_nchw().double()
    ref = act_ref(ref, act, 0.2)
    got = out.float_nchw().double()
    scale = ref.abs().max().item()
    assert torch.isfinite(got).all()
    if with_f32:
        assert (out_f32.permute(0, 3, 1, 2).double() - ref).abs().max().item() < 2e-5 * scale
    assert (got - ref).abs().max().item() < 3e-5 * scale  # + split-bf16 storage rounding (2^-17)


# mixed-kernel segment lists: (C, ksize, stride, pad) per segment
MIXED = [
    (2, 24, 32, [(64, 3, 1, 1), (64, 1, 1, 0), (128, 1, 1, 0)], 64, "lrelu"),   # BasicBlock conv2 + 1x1 shortcut segments
    (4, 96, 128, [(64, 3, 1, 1), (64, 3, 1, 1), (64, 3, 1, 1)], 64, "lrelu"),   # >= 148 items: M=256 halo tiles
    (2, 48, 64, [(128, 3, 1, 1), (128, 3, 2, 1)], 128, "lrelu"),                # never: mixed strides give different sizes
    (3, 26, 34, [(128, 3, 1, 0)], 16, "none"),                                  # valid conv over a pre-padded tensor
    (1, 48, 64, [(64, 3, 1, 1), (128, 3, 2, 1)], 128, "lrelu"),                 # stride-2 block: conv2 + 3x3/2 shortcut
]


@pytest.mark.parametrize("idx", [0, 1, 3, 4])
def test_conv_mixed_segments(idx):
    B, H, W, segs, Cout, act = MIXED[idx]
    torch.manual_seed(100 + idx)
    # segment 0 defines the output size; stride-2 segments read a 2x larger input
    C0, k0, s0, p0 = segs[0]
    OH = (H + 2 * p0 - k0) // s0 + 1
    OW = (W + 2 * p0 - k0) // s0 + 1
    xs, ws = [], []
    for (C, k, st, p) in segs:
        h_in = (OH - 1) * st + k - 2 * p
        w_in = (OW - 1) * st + k - 2 * p
        if st == 2:  # any size giving the same output works; use the even one
            h_in, w_in = OH * 2, OW * 2
        xs.append(torch.randn(B, C, h_in, w_in, device="cuda"))
        ws.append(torch.randn(Cout, C, k, k, device="cuda") / (C * k * k) ** 0.5)
    bias = torch.randn(Cout, device="cuda")
    acts = [SplitAct.from_nchw_torch(x) for x in xs]
    out = SplitAct(B, OH, OW, Cout, "cuda")
    plan = ConvPlan([(a, k, st, p) for a, (_, k, st, p) in zip(acts, segs)], ws, bias, out, B, Cout, act=act,
                    slope=0.2)
    plan.run()
    torch.cuda.synchronize()
    ref = sum(F.conv2d(a.float_nchw().double(), w.double(), None, st, p) for a, w, (_, k, st, p) in zip(acts, ws, segs))
    ref = act_ref(ref + bias.double().view(1, -1, 1, 1), act, 0.2)
    got = out.float_nchw().double()
    assert (got - ref).abs().max().item() < 3e-5 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 50, 70), (1, 96, 128), (3, 33, 47)])
def test_stem_conv7_tensor_core_matches_fp64(shape):
    """csrc/stem_tc.cu (7x7/2 stem as an implicit GEMM built row by row in tensor memory) vs torch fp64, incl. ragged
    sizes whose last tiles are clipped by the TMA store."""
    from implicit_depth_b200 import _abi
    from implicit_depth_b200.networks import Plan, ResnetMatchingEncoder

    n, H, W = shape
    torch.manual_seed(n * H + W)
    enc = ResnetMatchingEncoder(18, 16).cuda().eval()
    with torch.no_grad():
        enc.net[1].running_mean.normal_(0, 0.1)
        enc.net[1].running_var.uniform_(0.5, 1.5)
        enc.net[1].weight.uniform_(0.5, 1.5)
        enc.net[1].bias.normal_(0, 0.1)
    img = torch.randn(n, 3, H, W, device="cuda")
    # run only the first op of the plan (the stem) and read its output buffer back
    g = Plan("cuda")
    enc.plan(g, lambda: img, n, H, W)
    stem_op = g.ops[0][0]  # ops are (fn, reads, writes)
    stem_op()
    torch.cuda.synchronize()
    s1 = [c.cell_contents for c in stem_op.__closure__]
    act = [c for c in s1 if isinstance(c, SplitAct)][0]
    got = act.float_nchw().double()
    bn = enc.net[1]
    ref = F.conv2d(img.double(), enc.net[0].weight.double(), None, 2, 3)
    ref = F.batch_norm(ref, bn.running_mean.double(), bn.running_var.double(), bn.weight.double(), bn.bias.double(),
                       False, 0.0, bn.eps)
    ref = F.relu(ref)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 3e-5 * ref.abs().max().item()


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
@pytest.mark.parametrize("shape", [(2, 7, 9, 16), (1, 24, 32, 64), (3, 1, 5, 8), (1, 13, 1, 24)])
def test_upsample2x_matches_torch(shape, mode):
    """csrc/elementwise.cu upsample2x_kernel (one thread = a 2x2 output block from the 3x3 input neighbourhood) vs
    F.interpolate(scale_factor=2; bilinear with align_corners=False as utils/generic_utils.py:94-103, or nearest),
    incl. one-pixel-wide / -high maps where every tap clamps."""
    from implicit_depth_b200 import _abi

    B, H, W, C = shape
    torch.manual_seed(B * 1000 + H * 10 + W)
    a = SplitAct.from_nchw_torch(torch.randn(B, C, H, W, device="cuda"))
    out = SplitAct(B, 2 * H, 2 * W, C, "cuda")
    out.hi.fill_(float("nan")); out.lo.fill_(float("nan"))
    _abi.call("b200_upsample2x", _abi.ptr(a.hi), _abi.ptr(a.lo), _abi.ptr(out.hi), _abi.ptr(out.lo), B, H, W, C,
              0 if mode == "bilinear" else 1, _abi.stream_ptr())
    torch.cuda.synchronize()
    x = a.float_nchw()
    ref = F.interpolate(x, scale_factor=2, mode=mode, align_corners=False) if mode == "bilinear" else \
        F.interpolate(x, scale_factor=2, mode=mode)
    got = out.float_nchw()
    assert torch.isfinite(got).all()
    if mode == "nearest":
        assert torch.equal(got, ref)
    else:  # + the split-bf16 storage rounding of the output (2^-17 relative) and FMA contraction
        assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 48, 64, 64), (1, 31, 45, 16), (3, 20, 18, 64), (1, 192, 256, 64), (1, 5, 7, 8)])
def test_maxblurpool_matches_torch(shape):
    """MaxPool2d(2, stride 1) + BlurPool(filt 4, stride 2, reflect pad (1,2,1,2)) of antialiased-cnns 0.3 (call site
    modules/networks.py:267): sliding interior kernel + direct border rows against torch ops in fp64."""
    from implicit_depth_b200 import _abi

    B, H, W, C = shape
    torch.manual_seed(H * W + C)
    x = torch.randn(B, C, H, W, device="cuda")
    a = SplitAct.from_nchw_torch(x)
    xd = a.float_nchw().double()
    m = F.max_pool2d(xd, 2, 1)
    f = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64, device="cuda")
    filt = (f[:, None] * f[None, :] / 64.0)[None, None].repeat(C, 1, 1, 1)
    ref = F.conv2d(F.pad(m, [1, 2, 1, 2], mode="reflect"), filt, stride=2, groups=C)
    OH, OW = ref.shape[-2:]
    out = SplitAct(B, OH, OW, C, "cuda")
    out.hi.fill_(float("nan")); out.lo.fill_(float("nan"))
    _abi.call("b200_maxblurpool", _abi.ptr(a.hi), _abi.ptr(a.lo), _abi.ptr(out.hi), _abi.ptr(out.lo), B, H, W, C,
              _abi.stream_ptr())
    torch.cuda.synchronize()
    got = out.float_nchw().double()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("shape", [(2, 24, 32, 24, 96), (1, 96, 128, 8, 32), (1, 12, 16, 64, 256)])
def test_conv_stride2_tf_same_padding(shape):
    """TF "SAME" padding (timm `Conv2dSame`, the reference's tf_efficientnetv2_s encoder, bd_model.py:46-51) of a
    stride-2 3x3 conv on an even-sized map: padding (0, 1) -- nothing at the top / left, one pixel at the bottom /
    right -- as a shifted TMA box origin; the out-of-bounds column / row is zero-filled by the TMA unit."""
    B, H, W, C, Cout = shape
    torch.manual_seed(H + C)
    x = torch.randn(B, C, H, W, device="cuda")
    w = torch.randn(Cout, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    bias = torch.randn(Cout, device="cuda")
    a = SplitAct.from_nchw_torch(x)
    out = SplitAct(B, H // 2, W // 2, Cout, "cuda")
    plan = ConvPlan([(a, 3, 2, (0, 1))], [w], bias, out, B, Cout, act="silu")
    assert (plan.OH, plan.OW) == (H // 2, W // 2)
    plan.run()
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(F.pad(a.float_nchw().double(), [0, 1, 0, 1]), w.double(), bias.double(), 2, 0))
    assert (out.float_nchw().double() - ref).abs().max().item() < 3e-5 * ref.abs().max().item()
