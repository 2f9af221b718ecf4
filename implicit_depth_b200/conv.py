"""Host side of the tensor-core convolution path: split-bf16 NHWC activations, weight packing and
plan objects for `b200_conv_*` (csrc/conv_tc.cu)."""
from __future__ import annotations

import ctypes

import torch

from . import _abi

ACT = {"none": 0, "lrelu": 1, "elu": 2, "relu": 3, "silu": 4}
MAX_SEG = 4


class _Seg(ctypes.Structure):
    _fields_ = [("in_hi", ctypes.c_void_p), ("in_lo", ctypes.c_void_p), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("C", ctypes.c_int), ("ksize", ctypes.c_int), ("stride", ctypes.c_int), ("pad", ctypes.c_int),
                ("pad_hi", ctypes.c_int)]


class _Desc(ctypes.Structure):
    _fields_ = [("seg", _Seg * MAX_SEG), ("nseg", ctypes.c_int), ("wimage", ctypes.c_void_p),
                ("bias", ctypes.c_void_p), ("res_hi", ctypes.c_void_p), ("res_lo", ctypes.c_void_p),
                ("out_hi", ctypes.c_void_p), ("out_lo", ctypes.c_void_p), ("out_f32", ctypes.c_void_p),
                ("B", ctypes.c_int), ("OH", ctypes.c_int), ("OW", ctypes.c_int), ("Cout", ctypes.c_int),
                ("act", ctypes.c_int), ("slope", ctypes.c_float), ("max_ctas", ctypes.c_int),
                ("tile_hint", ctypes.c_int)]


class SplitAct:
    """An activation tensor in the tensor-core path's storage format: two NHWC bf16 planes with
    x = hi + lo (16 mantissa bits; same bytes as fp32)."""

    def __init__(self, B, H, W, C, device):
        self.B, self.H, self.W, self.C = B, H, W, C
        self.hi = torch.empty((B, H, W, C), device=device, dtype=torch.bfloat16)
        self.lo = torch.empty((B, H, W, C), device=device, dtype=torch.bfloat16)

    @property
    def shape(self):
        return (self.B, self.H, self.W, self.C)

    def float_nchw(self):
        """Debug/test helper (torch ops): reassemble an fp32 NCHW tensor."""
        return (self.hi.float() + self.lo.float()).permute(0, 3, 1, 2).contiguous()

    @staticmethod
    def from_nchw_torch(x):
        """Debug/test helper (torch ops); the product path uses b200_nchw_to_split."""
        B, C, H, W = x.shape
        a = SplitAct(B, H, W, C, x.device)
        xh = x.permute(0, 2, 3, 1).contiguous().float()
        a.hi.copy_(xh.to(torch.bfloat16))
        a.lo.copy_((xh - a.hi.float()).to(torch.bfloat16))
        return a


def pad_pair(p):
    """int (symmetric) or (lo, hi) -> (lo, hi)."""
    return (p, p) if isinstance(p, int) else (int(p[0]), int(p[1]))


def out_size(H, W, k, stride, pad):
    lo, hi = pad_pair(pad)
    return (H + lo + hi - k) // stride + 1, (W + lo + hi - k) // stride + 1


def same_pad(H, W, k, stride):
    """TF "SAME" padding (timm `Conv2dSame`) of a k x k conv: total = max((ceil(n / s) - 1) s + k - n, 0) per axis,
    split (total // 2, total - total // 2).  The kernels take one pair for both axes."""
    def one(n):
        total = max((-(-n // stride) - 1) * stride + k - n, 0)
        return total // 2, total - total // 2
    py, px = one(H), one(W)
    if py != px:
        raise NotImplementedError(f"'SAME' padding differs between the axes for a {H}x{W} map (stride {stride}): "
                                  "mixed even/odd sizes are not supported by the conv kernels")
    return py


def ntile(cout):
    return cout if cout <= 128 else 128


def _swizzle_last(t):
    """[..., R, 64] bf16 with R % 8 == 0 -> same shape, 16-byte chunk j of row r moved to chunk j ^ (r & 7)."""
    R = t.shape[-2]
    v = t.reshape(*t.shape[:-1], 8, 8)
    r = torch.arange(R, device=t.device).view(R, 1, 1)
    j = torch.arange(8, device=t.device).view(1, 8, 1)
    idx = (j ^ (r & 7)).expand(R, 8, 8).expand(v.shape)
    return torch.gather(v, -2, idx).reshape(t.shape)


def _swizzle64_last(t):
    """[..., R, 32] bf16 with R % 8 == 0 -> same shape, 16-byte chunk j of row r moved to chunk j ^ ((r >> 1) & 3)
    (the 64-byte swizzle of a dense tile whose rows are 64 B apart)."""
    R = t.shape[-2]
    v = t.reshape(*t.shape[:-1], 4, 8)
    r = torch.arange(R, device=t.device).view(R, 1, 1)
    j = torch.arange(4, device=t.device).view(1, 4, 1)
    idx = (j ^ ((r >> 1) & 3)).expand(R, 4, 8).expand(v.shape)
    return torch.gather(v, -2, idx).reshape(t.shape)


def pack_conv_weights(weights, seg_channels, cout, halo=False, nt=None):
    """weights: list (one per segment) of [Cout, C_s, k, k] fp32 tensors.  Returns the uint8 weight image in the
    producer's chunk order (plain: segment, tap, channel block; halo: segment, channel block, tap):
      plain kernel  [n_ntiles][chunk][hi NT x 64 | lo NT x 64]  64-channel chunks, 128-byte swizzle
      halo kernel   [n_ntiles][chunk][hi NT x 32 | lo NT x 32]  32-channel chunks, 64-byte swizzle; hi and lo are
                    adjacent rows of ONE tile so that A_hi x [B_hi | B_lo] is a single N = 2*NT MMA."""
    NT = ntile(cout) if nt is None else nt
    n_nt = (cout + NT - 1) // NT
    cw = 32 if halo else 64
    chunks = []
    for w, C in zip(weights, seg_channels):
        k = w.shape[-1]
        assert w.shape[0] == cout and w.shape[1] == C and w.shape[2] == k
        Cp = (C + cw - 1) // cw * cw
        wp = torch.zeros((cout, Cp, k, k), device=w.device, dtype=torch.float32)
        wp[:, :C] = w.detach().float()
        v = wp.view(cout, Cp // cw, cw, k, k)
        if halo:  # chunk order (channel block, tap): the 9 taps of a block are contiguous (one bulk copy per row)
            chunks.append(v.permute(1, 3, 4, 0, 2).reshape(-1, cout, cw))
        else:     # chunk order (tap, channel block)
            chunks.append(v.permute(3, 4, 1, 0, 2).reshape(-1, cout, cw))
    allw = torch.cat(chunks, 0)  # [Q, Cout, cw]
    Q = allw.shape[0]
    hi = allw.to(torch.bfloat16)
    lo = (allw - hi.float()).to(torch.bfloat16)
    both = torch.stack([hi, lo], 1).view(Q, 2, n_nt, NT, cw)  # [Q, part, nt, NT, cw]
    both = (_swizzle64_last(both) if halo else _swizzle_last(both)).permute(2, 0, 1, 3, 4).contiguous()
    return both.view(torch.uint8).reshape(-1)  # [nt, Q, part, NT, cw]


class ConvPlan:
    """One convolution launch with everything (tensor maps, weight image, buffers) fixed at plan time."""

    def __init__(self, segs, weights, bias, out, B, cout, act="none", slope=0.2, residual=None, out_f32=None,
                 max_ctas=0, tile_hint=0):
        """segs: list of (SplitAct, ksize, stride, pad); `pad` is an int (torch-style symmetric padding) or a pair
        (top/left, bottom/right) -- (0, 1) is the TF "SAME" padding of a stride-2 3x3 conv on an even-sized map;
        weights: one [Cout, C, k, k] fp32 tensor per segment; out: SplitAct or None; residual: SplitAct or None;
        max_ctas: CTA cap of the persistent launch (0 = all SMs)."""
        d = _Desc()
        d.nseg = len(segs)
        for i, (a, k, s, p) in enumerate(segs):
            p_lo, p_hi = pad_pair(p)
            d.seg[i].in_hi = a.hi.data_ptr()
            d.seg[i].in_lo = a.lo.data_ptr()
            d.seg[i].H, d.seg[i].W, d.seg[i].C = a.H, a.W, a.C
            d.seg[i].ksize, d.seg[i].stride, d.seg[i].pad, d.seg[i].pad_hi = k, s, p_lo, p_hi
        a0, k0, s0, p0 = segs[0]
        OH, OW = out_size(a0.H, a0.W, k0, s0, p0)
        d.bias = bias.data_ptr() if bias is not None else None
        d.res_hi = residual.hi.data_ptr() if residual is not None else None
        d.res_lo = residual.lo.data_ptr() if residual is not None else None
        d.out_hi = out.hi.data_ptr() if out is not None else None
        d.out_lo = out.lo.data_ptr() if out is not None else None
        d.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
        d.B, d.OH, d.OW, d.Cout = B, OH, OW, cout
        d.act = ACT[act]
        d.slope = slope
        d.max_ctas = int(max_ctas)
        d.tile_hint = int(tile_hint)
        lib = _abi.load()
        self.halo = bool(lib.b200_conv_uses_halo(ctypes.byref(d)))  # decides the weight-image layout
        self.nt = int(lib.b200_conv_ntile_for(ctypes.byref(d)))  # N tile of this conv (weight-image layout)
        wimage = pack_conv_weights(weights, [a.C for a, _, _, _ in segs], cout, halo=self.halo, nt=self.nt)
        d.wimage = wimage.data_ptr()
        self._keep = (segs, wimage, bias, out, residual, out_f32)  # keep buffers alive
        self.handle = ctypes.c_void_p()
        rc = lib.b200_conv_create(ctypes.byref(d), ctypes.byref(self.handle))
        if rc != 0:
            raise _abi.B200Error(f"b200_conv_create failed ({rc}): {lib.b200_last_error().decode()}")
        self.OH, self.OW = OH, OW

    def run(self):
        _abi.call("b200_conv_run", self.handle, _abi.stream_ptr())

    def __del__(self):
        try:
            if self.handle:
                _abi.load().b200_conv_destroy(self.handle)
        except Exception:
            pass
