"""B200 versions of the reference's convolutional networks.

Each class keeps the reference module's constructor arguments and state-dict keys (so its checkpoints
load) but holds no PyTorch compute: `forward` builds -- once per input shape -- a static launch plan of
hand-written sm_100a kernels (tcgen05 implicit-GEMM convs over NHWC split-bf16 activations, plus the
layout / upsample / instance-norm kernels) and replays it.

  ResnetMatchingEncoder  modules/networks.py:236-287   (antialiased ResNet-18 stem + IN head)
  CVEncoder              modules/networks.py:186-215
  BDDecoderPP            modules/networks.py:20-84     (UNet++)
  SkipDecoder            modules/networks_fast.py:49-99
  BinaryMLPNetwork       modules/networks.py:87-115    (+ BDModel.run_mlp_val, bd_model.py:412-442)
  BasicBlock             modules/layers.py:34-95
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch import nn

from . import _abi
from .conv import ConvPlan, SplitAct

LRELU = 0.2  # nn.LeakyReLU(0.2) of BasicBlock (layers.py:60) and the matching head (networks.py:278)


# =====================================================================================
# static launch plan
# =====================================================================================
class Plan:
    """Preallocated buffers + an ordered list of kernel launches (CUDA-graph capturable)."""

    DAG = True       # False = strictly sequential launches on the caller's stream (the bit-identity tests compare both)
    N_STREAMS = 5    # extra streams of the dependency scheduler (scripts/stream_sweep.py, ms per step at cfg2: 1 -> 6.92,
                     # 2 -> 6.83, 3 -> 6.79, 4 -> 6.78, 5 -> 6.68, 6 -> 6.73, 8 -> 6.80; profiles/r02x_stream_sweep.jsonl)

    def __init__(self, device, max_ctas=0):
        self.device = torch.device(device)
        self.ops = []  # (fn, reads | None, writes | None)
        self.n_launches = 0
        self._keep = []
        self.dag = Plan.DAG
        self.n_streams = Plan.N_STREAMS
        self.max_ctas = int(max_ctas)  # CTA cap of this plan's persistent kernels (0 = all SMs)
        self._pool = None
        self.external = {}  # id(buffer produced outside this plan, possibly on another stream) -> event to wait on

    def act(self, B, H, W, C):
        return SplitAct(B, H, W, C, self.device)

    def empty(self, shape, dtype=torch.float32):
        t = torch.empty(shape, device=self.device, dtype=dtype)
        self._keep.append(t)
        return t

    def add(self, fn, launches=1, reads=None, writes=None):
        """Append one op.  `reads` / `writes`: the buffers (SplitActs / tensors of this plan) it touches -- ops that
        declare them are scheduled by data dependence (see `run`); an op that declares nothing is a barrier: it
        runs on the caller's stream after everything before it and before everything after it."""
        self.ops.append((fn, None if reads is None else [b for b in reads if b is not None],
                         None if writes is None else [b for b in writes if b is not None]))
        self.n_launches += launches

    def run(self):
        """Launch the plan.  Ops with declared buffers are list-scheduled over a few CUDA streams: an op goes onto
        the stream whose tail produced one of its inputs (chains stay on one stream, no event needed), otherwise onto
        an idle / round-robin stream, and waits on the events of its other producers (and of earlier readers of a
        buffer it overwrites).  Under CUDA-graph capture this becomes the true dependency DAG of the step, so the
        independent BasicBlocks of the UNet++ decoder and the cost-volume encoder chain run side by side: kernels of
        small grids (low-resolution layers) share the machine and the tails of full-grid kernels overlap."""
        if not self.dag or all(r is None for _, r, _ in self.ops):
            for ev in self.external.values():
                torch.cuda.current_stream().wait_event(ev)
            for fn, _, _ in self.ops:
                fn()
            return
        main = torch.cuda.current_stream()
        ext_waited = set()
        if self._pool is None:
            self._pool = [_abi.new_stream(self.device) for _ in range(self.n_streams)]
        streams = [main] + self._pool
        n = len(streams)
        tail = [None] * n            # index of the last op enqueued on each stream since the last barrier
        synced = [True] + [False] * (n - 1)
        start = torch.cuda.Event()
        start.record(main)
        last_write, readers, op_stream, op_event = {}, {}, {}, {}
        rr = 0
        for idx, (fn, reads, writes) in enumerate(self.ops):
            if reads is None:  # barrier
                for k in range(1, n):
                    if tail[k] is not None:
                        main.wait_stream(streams[k])
                for b, ev in self.external.items():
                    if (0, b) not in ext_waited:
                        main.wait_event(ev)
                        ext_waited.add((0, b))
                fn()
                tail = [None] * n
                synced = [True] + [False] * (n - 1)
                start = torch.cuda.Event()
                start.record(main)
                last_write, readers = {}, {}
                continue
            deps = set()
            for b in reads:
                if id(b) in last_write:
                    deps.add(last_write[id(b)])
            for b in writes:
                if id(b) in last_write:
                    deps.add(last_write[id(b)])
                deps.update(readers.get(id(b), ()))
            k = None
            for d in sorted(deps, reverse=True):
                if tail[op_stream[d]] == d:
                    k = op_stream[d]
                    break
            if k is None:
                idle = [q for q in range(n) if tail[q] is None]
                if idle:
                    k = idle[0]
                else:
                    rr = rr % n
                    k = rr
                    rr += 1
            st = streams[k]
            if not synced[k]:
                st.wait_event(start)
                synced[k] = True
            for d in deps:
                if op_stream[d] != k:
                    st.wait_event(op_event[d])
            for b in reads:
                if id(b) in self.external and (k, id(b)) not in ext_waited:
                    st.wait_event(self.external[id(b)])
                    ext_waited.add((k, id(b)))
            if k == 0:
                fn()
            else:
                with torch.cuda.stream(st):
                    fn()
            ev = torch.cuda.Event()
            ev.record(st)
            op_event[idx], op_stream[idx], tail[k] = ev, k, idx
            for b in reads:
                readers.setdefault(id(b), []).append(idx)
            for b in writes:
                last_write[id(b)] = idx
                readers[id(b)] = []
        for k in range(1, n):
            if tail[k] is not None:
                main.wait_stream(streams[k])

    # ---- op builders --------------------------------------------------------------------
    def conv(self, segs, bias, cout, act="none", slope=LRELU, residual=None, want_f32=False, want_split=True,
             out=None, out_f32=None):
        """segs: list of (SplitAct, weight [Cout, C, k, k], stride, pad).  Returns (SplitAct | None, f32 | None).
        `out` / `out_f32` may name existing buffers to overwrite (re-used scratch)."""
        a0, w0, s0, p0 = segs[0]
        k0 = w0.shape[-1]
        B = a0.B
        from .conv import out_size

        OH, OW = out_size(a0.H, a0.W, k0, s0, p0)
        if out is None and want_split:
            out = self.act(B, OH, OW, cout)
        if out_f32 is None and want_f32:
            out_f32 = self.empty((B, OH, OW, cout))
        b = None if bias is None else bias.detach().to(self.device, torch.float32).contiguous()
        plan = ConvPlan([(a, w.shape[-1], s, p) for a, w, s, p in segs], [w for _, w, _, _ in segs], b, out, B, cout, act=act,
                        slope=slope, residual=residual, out_f32=out_f32, max_ctas=self.max_ctas)
        self.add(plan.run, reads=[a for a, _, _, _ in segs] + [residual], writes=[out, out_f32])
        return out, out_f32

    def upsample2x(self, a, mode):
        out = self.act(a.B, 2 * a.H, 2 * a.W, a.C)
        m = {"bilinear": 0, "nearest": 1}[mode]
        self.add(lambda: _abi.call("b200_upsample2x", _abi.ptr(a.hi), _abi.ptr(a.lo), _abi.ptr(out.hi),
                                   _abi.ptr(out.lo), a.B, a.H, a.W, a.C, m, _abi.stream_ptr()), reads=[a], writes=[out])
        return out

    def from_f32(self, getter, B, C, H, W):
        """fp32 [B,C,H,W] tensor (any strides, fetched at run time through `getter`) -> SplitAct."""
        out = self.act(B, H, W, C)

        def op():
            t = getter()
            if t.dtype != torch.float32:
                t = t.float()
            assert tuple(t.shape) == (B, C, H, W), f"expected {(B, C, H, W)}, got {tuple(t.shape)}"
            _abi.require_cuda(t)
            _abi.call("b200_f32_to_split", _abi.ptr(t), _abi.ptr(out.hi), _abi.ptr(out.lo), B, C, H, W, t.stride(0),
                      t.stride(1), t.stride(2), t.stride(3), _abi.stream_ptr())

        self.add(op, reads=[], writes=[out])
        return out

    def to_nchw(self, a):
        out = self.empty((a.B, a.C, a.H, a.W))
        self.add(lambda: _abi.call("b200_split_to_nchw", _abi.ptr(a.hi), _abi.ptr(a.lo), _abi.ptr(out), a.B, a.C, a.H,
                                   a.W, _abi.stream_ptr()), reads=[a], writes=[out])
        return out

    def channel_dot(self, a, conv):
        """nn.Conv2d(C, 1, 1) on a split activation -> (log fp32 [B,1,H,W], exp(log) fp32 [B,1,H,W])."""
        assert conv.out_channels == 1 and conv.kernel_size == (1, 1)
        cl = conv.in_channels
        w = torch.zeros(a.C, device=self.device, dtype=torch.float32)
        w[:cl] = conv.weight.detach().to(self.device, torch.float32).reshape(-1)
        b = conv.bias.detach().to(self.device, torch.float32).reshape(1).contiguous()
        out_log = self.empty((a.B, 1, a.H, a.W))
        out_exp = self.empty((a.B, 1, a.H, a.W))
        self._keep += [w, b]
        self.add(lambda: _abi.call("b200_channel_dot_exp", _abi.ptr(a.hi), _abi.ptr(a.lo), _abi.ptr(w), _abi.ptr(b),
                                   _abi.ptr(out_log), _abi.ptr(out_exp), a.B * a.H * a.W, a.C, _abi.stream_ptr()),
                 reads=[a], writes=[out_log, out_exp])
        return out_log, out_exp

    def instance_norm(self, a, pad=0, act="none", slope=LRELU, eps=1e-5, f32_pixel_major=False, f32_layout=0):
        partial = self.empty((a.B, 32, a.C, 2), torch.float64)
        stats = self.empty((a.B, a.C, 2))
        out = None if f32_pixel_major else self.act(a.B, a.H + 2 * pad, a.W + 2 * pad, a.C)
        out32 = self.empty((a.B, a.H * a.W, a.C)) if f32_pixel_major else None
        self.add(lambda: _abi.call("b200_instance_norm", _abi.ptr(a.hi), _abi.ptr(a.lo), _abi.ptr(partial),
                                   _abi.ptr(stats), _abi.ptr(out.hi) if out else None,
                                   _abi.ptr(out.lo) if out else None, _abi.ptr(out32), a.B, a.H, a.W, a.C, pad,
                                   1 if act == "lrelu" else 0, slope, eps, f32_layout, _abi.stream_ptr()),
                 launches=3)
        return out if out is not None else out32


def _split_weight(w, parts):
    """Split a conv weight [Cout, sum(C_i), k, k] along its input channels like torch.cat's inputs."""
    sizes = [getattr(p, "Cl", p.C) for p in parts]  # logical channels (zero-padded activations carry p.C > p.Cl)
    assert sum(sizes) == w.shape[1], f"concat widths {sizes} do not match conv input {w.shape[1]}"
    pieces = list(torch.split(w.detach(), sizes, dim=1))
    for i, (p, piece) in enumerate(zip(parts, pieces)):
        if p.C != piece.shape[1]:  # zero weight columns for the padding channels
            pieces[i] = torch.cat([piece, piece.new_zeros((piece.shape[0], p.C - piece.shape[1]) + tuple(piece.shape[2:]))], 1)
    return pieces


# =====================================================================================
# parameter containers (state-dict compatible with the reference) + plan builders
# =====================================================================================
class BasicBlock(nn.Module):
    """Residual block of modules/layers.py:34-95 with norm_layer=nn.Identity (biased convs):
    conv3x3 -> LeakyReLU(0.2) -> conv3x3 -> (+ identity | 1x1 conv | 3x3 stride-2 conv) -> LeakyReLU(0.2)."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=True)
        if inplanes == planes and stride == 1:
            self.downsample = None
        else:
            k = 1 if stride == 1 else 3
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, k, stride, k // 2, bias=True), nn.Identity())
        self.stride = stride

    def plan(self, g: Plan, parts, act="lrelu", want_f32=False):
        """parts: list of SplitActs that the reference would torch.cat along channels."""
        planes = self.conv1.out_channels
        w1 = _split_weight(self.conv1.weight, parts)
        h, _ = g.conv([(p, w, self.stride, 1) for p, w in zip(parts, w1)], self.conv1.bias, planes, act=act)
        if self.downsample is None:
            assert len(parts) == 1
            return g.conv([(h, self.conv2.weight, 1, 1)], self.conv2.bias, planes, act=act, residual=parts[0],
                          want_f32=want_f32)
        ds = self.downsample[0]
        k = ds.kernel_size[0]
        wd = _split_weight(ds.weight, parts)
        segs = [(h, self.conv2.weight, 1, 1)] + [(p, w, self.stride, k // 2) for p, w in zip(parts, wd)]
        return g.conv(segs, self.conv2.bias.detach() + ds.bias.detach(), planes, act=act, want_f32=want_f32)


class _PlannedModule(nn.Module):
    """Caches one launch plan per input signature; invalidated when parameters change."""

    def __init__(self):
        super().__init__()
        self._plans = {}

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple(
            (b.data_ptr(), b._version) for b in self.buffers())

    def _get_plan(self, sig, build):
        key = (sig, self._param_key())
        if key not in self._plans:
            self._plans = {key: build()}  # keep only the latest signature
        return self._plans[key]

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)


class CVEncoder(_PlannedModule):
    """modules/networks.py:186-215: per level BasicBlock(stride 1|2) -> cat image feature -> 2 BasicBlocks."""

    def __init__(self, num_ch_cv, num_ch_enc, num_ch_outs):
        super().__init__()
        self.convs = nn.ModuleDict()
        self.num_ch_enc = []
        self.num_blocks = len(num_ch_outs)
        for i in range(self.num_blocks):
            cin = num_ch_cv if i == 0 else num_ch_outs[i - 1]
            cout = num_ch_outs[i]
            self.convs[f"ds_conv_{i}"] = BasicBlock(cin, cout, stride=1 if i == 0 else 2)
            self.convs[f"conv_{i}"] = nn.Sequential(BasicBlock(num_ch_enc[i] + cout, cout), BasicBlock(cout, cout))
            self.num_ch_enc.append(cout)

    def plan(self, g: Plan, x, img_feats):
        outs = []
        for i in range(self.num_blocks):
            x, _ = self.convs[f"ds_conv_{i}"].plan(g, [x])
            x, _ = self.convs[f"conv_{i}"][0].plan(g, [x, img_feats[i]])
            x, _ = self.convs[f"conv_{i}"][1].plan(g, [x])
            outs.append(x)
        return outs

    @torch.no_grad()
    def forward(self, x, img_feats):
        sig = (tuple(x.shape),) + tuple(tuple(f.shape) for f in img_feats)

        def build():
            g = Plan(x.device)
            slots = {}
            xa = g.from_f32(lambda: slots["x"], *x.shape)
            fa = [g.from_f32((lambda i=i: slots[f"f{i}"]), *f.shape) for i, f in enumerate(img_feats)]
            outs = [g.to_nchw(o) for o in self.plan(g, xa, fa)]
            return g, slots, outs

        g, slots, outs = self._get_plan(sig, build)
        slots["x"] = x
        for i, f in enumerate(img_feats):
            slots[f"f{i}"] = f
        g.run()
        return [o.clone() for o in outs]


def _double_basic_block(cin, cout):
    seq = nn.Sequential(BasicBlock(cin, cout))
    seq.add_module("conv_0", BasicBlock(cout, cout))  # key layout of networks.py:13-17
    return seq


class BDDecoderPP(_PlannedModule):
    """UNet++ decoder of modules/networks.py:20-84.  `outputs` selects which feature_s{i}_b1hw maps are
    produced; the BD inference path reads only feature_s0 (bd_model.py:414), whose output_0 is the identity,
    so with outputs=(0,) every output_* BasicBlock (dead work in the reference) is skipped."""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=16, use_skips=True, outputs=(0, 1, 2, 3)):
        super().__init__()
        self.num_ch_enc = list(num_ch_enc)
        self.num_ch_dec = np.array([64, 64, 128, 256])
        self.outputs = tuple(outputs)
        self.convs = nn.ModuleDict()
        dec = [int(c) for c in self.num_ch_dec]
        for j in range(1, 5):
            for i in range(4 - j, -1, -1):
                cout = dec[i]
                total = 0
                self.convs[f"diag_conv_{i + 1}{j - 1}"] = BasicBlock(self.num_ch_enc[i + 1] if j == 1 else dec[i + 1], cout)
                total += cout
                self.convs[f"right_conv_{i}{j - 1}"] = BasicBlock(self.num_ch_enc[i] if j == 1 else dec[i], cout)
                total += cout
                if i + j != 4:
                    self.convs[f"up_conv_{i + 1}{j}"] = BasicBlock(dec[i + 1], cout)
                    total += cout
                self.convs[f"in_conv_{i}{j}"] = _double_basic_block(total, cout)
                self.convs[f"output_{i}"] = nn.Sequential(BasicBlock(cout, cout) if i != 0 else nn.Identity())

    def plan(self, g: Plan, feats, outputs=None, want_f32=()):
        """feats: 5 SplitActs (/2 ... /32).  Returns {i: SplitAct} (+ {i: fp32 NHWC} for i in want_f32)."""
        outputs = self.outputs if outputs is None else tuple(outputs)
        prev = list(feats)
        outs = []
        result, result_f32 = {}, {}
        for j in range(1, 5):
            for i in range(4 - j, -1, -1):
                parts = [self.convs[f"right_conv_{i}{j - 1}"].plan(g, [prev[i]])[0]]
                d, _ = self.convs[f"diag_conv_{i + 1}{j - 1}"].plan(g, [prev[i + 1]])
                parts.append(g.upsample2x(d, "bilinear"))
                if i + j != 4:
                    u, _ = self.convs[f"up_conv_{i + 1}{j}"].plan(g, [outs[-1]])
                    parts.append(g.upsample2x(u, "bilinear"))
                blk = self.convs[f"in_conv_{i}{j}"]
                x, _ = blk[0].plan(g, parts)
                last = (j == 4 - i) and (i in outputs)
                is_final_feature = last and i == 0  # output_0 is the identity
                x, x32 = blk[1].plan(g, [x], want_f32=is_final_feature and (0 in want_f32))
                outs.append(x)
                if last:
                    if i == 0:
                        result[0] = x
                        if x32 is not None:
                            result_f32[0] = x32
                    else:
                        o, o32 = self.convs[f"output_{i}"][0].plan(g, [x], want_f32=i in want_f32)
                        result[i] = o
                        if o32 is not None:
                            result_f32[i] = o32
            prev = outs[::-1]
        return result, result_f32

    @torch.no_grad()
    def forward(self, input_features):
        sig = tuple(tuple(f.shape) for f in input_features)

        def build():
            g = Plan(input_features[0].device)
            slots = {}
            fa = [g.from_f32((lambda i=i: slots[i]), *f.shape) for i, f in enumerate(input_features)]
            res, _ = self.plan(g, fa)
            return g, slots, {i: g.to_nchw(a) for i, a in res.items()}

        g, slots, outs = self._get_plan(sig, build)
        for i, f in enumerate(input_features):
            slots[i] = f
        g.run()
        return {f"feature_s{i}_b1hw": o.clone() for i, o in outs.items()}


class DepthDecoderPP(BDDecoderPP):
    """Regression UNet++ of modules/networks.py:118-183: the BD decoder's graph with every `output_i` live and a
    1x1 conv to one log-depth channel behind it (`convs.output_{i}.1`)."""

    def __init__(self, num_ch_enc, scales=range(4), num_output_channels=1, use_skips=True):
        super().__init__(num_ch_enc, outputs=(0, 1, 2, 3))
        if num_output_channels != 1:
            raise ValueError("B200 DepthDecoderPP predicts one log-depth channel")
        for i in range(4):
            c = int(self.num_ch_dec[i])
            self.convs[f"output_{i}"] = nn.Sequential(BasicBlock(c, c) if i != 0 else nn.Identity(), nn.Conv2d(c, 1, 1))

    def plan_depth(self, g: Plan, feats):
        """Returns {i: (log_depth, depth)} fp32 [B,1,H_i,W_i] for the four scales."""
        res, _ = self.plan(g, feats, outputs=(0, 1, 2, 3))
        return {i: g.channel_dot(res[i], self.convs[f"output_{i}"][1]) for i in range(4)}

    @torch.no_grad()
    def forward(self, input_features):
        sig = tuple(tuple(f.shape) for f in input_features)

        def build():
            g = Plan(input_features[0].device)
            slots = {}
            fa = [g.from_f32((lambda i=i: slots[i]), *f.shape) for i, f in enumerate(input_features)]
            return g, slots, self.plan_depth(g, fa)

        g, slots, outs = self._get_plan(sig, build)
        for i, f in enumerate(input_features):
            slots[i] = f
        g.run()
        return {f"log_depth_pred_s{i}_b1hw": o[0].clone() for i, o in outs.items()}


class _ConvBlock(nn.Module):
    """modules/networks_fast.py:10-28: conv3x3 -> ELU -> conv3x3 -> ELU."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)

    def plan(self, g, parts, want_f32=False):
        cout = self.conv1.out_channels
        w1 = _split_weight(self.conv1.weight, parts)
        h, _ = g.conv([(p, w, 1, 1) for p, w in zip(parts, w1)], self.conv1.bias, cout, act="elu")
        return g.conv([(h, self.conv2.weight, 1, 1)], self.conv2.bias, cout, act="elu", want_f32=want_f32)


class _UpConcatBlock(nn.Module):
    """modules/networks_fast.py:31-46."""

    def __init__(self, cin, cout, skip):
        super().__init__()
        self.pre_concat_conv = _ConvBlock(cin, cout)
        self.post_concat_conv = _ConvBlock(cout + skip, cout)

    def plan(self, g, x, skip, want_f32=False):
        x, _ = self.pre_concat_conv.plan(g, [x])
        x = g.upsample2x(x, "nearest")
        return self.post_concat_conv.plan(g, [x, skip], want_f32=want_f32)


class SkipDecoder(_PlannedModule):
    """Plain U-Net decoder of modules/networks_fast.py:49-99."""

    def __init__(self, input_channels, use_bn=False):
        super().__init__()
        ic = list(input_channels)[::-1]
        self.input_channels = ic
        self.output_channels = [256, 128, 64, 64]
        self.num_ch_dec = self.output_channels[::-1]
        for n in range(4):
            setattr(self, f"block{n + 1}", _UpConcatBlock(ic[n], self.output_channels[n], ic[n + 1]))

    def plan(self, g, feats, outputs=(0, 1, 2, 3), want_f32=()):
        x = feats[-1]
        result, result_f32 = {}, {}
        for n in range(4):
            scale = 3 - n
            x, x32 = getattr(self, f"block{n + 1}").plan(g, x, feats[-2 - n], want_f32=scale in want_f32)
            if scale in outputs:
                result[scale] = x
                if x32 is not None:
                    result_f32[scale] = x32
        return result, result_f32

    @torch.no_grad()
    def forward(self, features):
        sig = tuple(tuple(f.shape) for f in features)

        def build():
            g = Plan(features[0].device)
            slots = {}
            fa = [g.from_f32((lambda i=i: slots[i]), *f.shape) for i, f in enumerate(features)]
            res, _ = self.plan(g, fa)
            return g, slots, {i: g.to_nchw(a) for i, a in res.items()}

        g, slots, outs = self._get_plan(sig, build)
        for i, f in enumerate(features):
            slots[i] = f
        g.run()
        return {f"feature_s{i}_b1hw": o.clone() for i, o in outs.items()}


class SkipDecoderRegression(SkipDecoder):
    """modules/networks_fast.py:102-145: SkipDecoder + one 1x1-conv MLP head (C -> 128 -> 128 -> 1, ELU) per scale;
    `out1` reads feature_s3 ... `out4` reads feature_s0."""

    def __init__(self, input_channels, use_bn=False):
        super().__init__(input_channels, use_bn=use_bn)
        for n in range(4):
            setattr(self, f"out{n + 1}", nn.Sequential(
                nn.Conv2d(self.output_channels[n], 128, 1), nn.ELU(inplace=True), nn.Conv2d(128, 128, 1),
                nn.ELU(inplace=True), nn.Conv2d(128, 1, 1)))

    def plan_depth(self, g, feats):
        res, _ = self.plan(g, feats, outputs=(0, 1, 2, 3))
        out = {}
        for n in range(4):
            scale = 3 - n
            head = getattr(self, f"out{n + 1}")
            h, _ = g.conv([(res[scale], head[0].weight, 1, 0)], head[0].bias, 128, act="elu")
            h, _ = g.conv([(h, head[2].weight, 1, 0)], head[2].bias, 128, act="elu")
            out[scale] = g.channel_dot(h, head[4])
        return out

    @torch.no_grad()
    def forward(self, features):
        sig = tuple(tuple(f.shape) for f in features)

        def build():
            g = Plan(features[0].device)
            slots = {}
            fa = [g.from_f32((lambda i=i: slots[i]), *f.shape) for i, f in enumerate(features)]
            return g, slots, self.plan_depth(g, fa)

        g, slots, outs = self._get_plan(sig, build)
        for i, f in enumerate(features):
            slots[i] = f
        g.run()
        return {f"log_depth_pred_s{i}_b1hw": o[0].clone() for i, o in outs.items()}


# ---- matching encoder ---------------------------------------------------------------
class _BlurPool(nn.Module):
    def __init__(self, channels):
        super().__init__()
        a = torch.tensor([1.0, 3.0, 3.0, 1.0])
        f = a[:, None] * a[None, :]
        self.register_buffer("filt", (f / f.sum())[None, None].repeat(channels, 1, 1, 1))


class _ResBlockBN(nn.Module):
    """torchvision-style BasicBlock(64, 64) with BatchNorm + ReLU (layer1 of the antialiased ResNet-18)."""

    def __init__(self, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)


def _fold_bn(conv_w, bn):
    """Eval-mode BatchNorm folded into the preceding bias-free conv."""
    s = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    return conv_w.detach() * s.view(-1, 1, 1, 1), bn.bias.detach() - bn.running_mean.detach() * s


class ResnetMatchingEncoder(_PlannedModule):
    """modules/networks.py:236-287 with num_layers=18, antialiased=True: `net` = [conv1, bn1, relu,
    maxpool(+BlurPool), layer1, 1x1 conv 64->128, InstanceNorm, LeakyReLU(0.2), 3x3 replicate-pad conv
    128->C, InstanceNorm].  Arithmetic of the third-party stem restated from antialiased-cnns 0.3
    (parity unpinned for that part).  Output: [n, C, H/4, W/4]."""

    def __init__(self, num_layers=18, num_ch_out=16, pretrained=False, antialiased=True):
        super().__init__()
        if num_layers != 18 or not antialiased:
            raise ValueError("B200 matching encoder implements the antialiased ResNet-18 stem only")
        self.num_ch_enc = np.array([64, 64])
        self.num_ch_out = num_ch_out
        self.net = nn.Sequential(
            nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False),
            nn.BatchNorm2d(64),
            nn.ReLU(inplace=True),
            nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=1), _BlurPool(64)),
            nn.Sequential(_ResBlockBN(64), _ResBlockBN(64)),
            nn.Conv2d(64, 128, (1, 1)),
            nn.InstanceNorm2d(128),
            nn.LeakyReLU(0.2, True),
            nn.Conv2d(128, num_ch_out, (3, 3), padding=1, padding_mode="replicate"),
            nn.InstanceNorm2d(num_ch_out),
        )

    def plan(self, g: Plan, get_images, n, H, W, feat_layout=0):
        """Returns the fp32 pixel-major feature buffer [n, (H/4)*(W/4), C] the volume kernels consume, in gather layout
        `feat_layout` (csrc/common.cuh: 0 = texel records, 1 = quarter-planar)."""
        net = self.net
        dev = g.device
        w7, b7 = _fold_bn(net[0].weight, net[1])
        # K index of the tensor-core stem kernel (csrc/stem_tc.cu): k = (c*7 + dy)*8 + dx, dx = 7 is a zero column
        wk = torch.zeros((64, 3, 7, 8), dtype=torch.float32)
        wk[..., :7] = w7.detach().float().cpu()
        wk = torch.cat([wk.reshape(64, 168), torch.zeros(64, 24)], 1).to(dev)  # [64, 192]
        hi = wk.to(torch.bfloat16)
        lo = (wk - hi.float()).to(torch.bfloat16)
        tiles = torch.cat([hi, lo], 0).view(128, 3, 64).permute(1, 0, 2).contiguous()  # [chunk, 128 rows, 64 k]
        from .conv import _swizzle_last

        wimage = _swizzle_last(tiles).contiguous().view(torch.uint8).reshape(-1)
        b7 = b7.to(dev, torch.float32).contiguous()
        H2, W2 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        s1 = g.act(n, H2, W2, 64)
        g._keep += [wimage, b7]

        def stem():
            img = get_images()
            assert tuple(img.shape) == (n, 3, H, W) and img.is_contiguous() and img.dtype == torch.float32
            # (not capped by g.max_ctas: the first kernel of the forward runs before the image encoder's stream has
            # anything to share the SMs with)
            _abi.call("b200_stem_conv7_tc", _abi.ptr(img), _abi.ptr(wimage), _abi.ptr(b7), _abi.ptr(s1.hi),
                      _abi.ptr(s1.lo), n, H, W, 0, _abi.stream_ptr())

        g.add(stem)
        H4, W4 = (H2 - 2) // 2 + 1, (W2 - 2) // 2 + 1
        x = g.act(n, H4, W4, 64)
        g.add(lambda o=x: _abi.call("b200_maxblurpool", _abi.ptr(s1.hi), _abi.ptr(s1.lo), _abi.ptr(o.hi),
                                    _abi.ptr(o.lo), n, H2, W2, 64, _abi.stream_ptr()))  # bind now: x is rebound below
        for blk in net[4]:
            w1, b1 = _fold_bn(blk.conv1.weight, blk.bn1)
            w2, b2 = _fold_bn(blk.conv2.weight, blk.bn2)
            h, _ = g.conv([(x, w1, 1, 1)], b1, 64, act="relu")
            x, _ = g.conv([(h, w2, 1, 1)], b2, 64, act="relu", residual=x)
        x, _ = g.conv([(x, net[5].weight, 1, 0)], net[5].bias, 128, act="none")
        x = g.instance_norm(x, pad=1, act="lrelu", slope=0.2, eps=net[6].eps)  # replicate border for net[8]
        x, _ = g.conv([(x, net[8].weight, 1, 0)], net[8].bias, self.num_ch_out, act="none")
        return g.instance_norm(x, eps=net[9].eps, f32_pixel_major=True, f32_layout=feat_layout), H4, W4

    @torch.no_grad()
    def forward(self, input_image):
        n, _, H, W = input_image.shape
        sig = tuple(input_image.shape)

        def build():
            g = Plan(input_image.device)
            slots = {}
            feats, H4, W4 = self.plan(g, lambda: slots["img"], n, H, W)
            return g, slots, feats, H4, W4

        g, slots, feats, H4, W4 = self._get_plan(sig, build)
        slots["img"] = input_image.contiguous().float()
        g.run()
        return feats.view(n, H4, W4, self.num_ch_out).permute(0, 3, 1, 2).contiguous()


# ---- binary occupancy MLP -------------------------------------------------------------
class BinaryMLPNetwork(nn.Module):
    """modules/networks.py:87-115: per scale Linear(C+extra,128) -> ELU -> Linear(128,128) -> ELU -> Linear(128,1).
    Inference uses scale 0 only (max_scale_only, bd_model.py:439)."""

    def __init__(self, input_channels, mlp_size=128, use_prior=False):
        super().__init__()
        self.scales = list(range(4))
        self.use_prior = use_prior
        extra = 2 if use_prior else 1
        self.mlps = nn.ModuleDict()
        for scale, c in enumerate(input_channels):
            self.mlps[f"s{scale}"] = nn.Sequential(nn.Linear(int(c) + extra, mlp_size), nn.ELU(inplace=True),
                                                   nn.Linear(mlp_size, mlp_size), nn.ELU(inplace=True),
                                                   nn.Linear(mlp_size, 1))

    def _pack(self, dev):
        """Shared-memory weight image + fp32 vectors of the fused kernel (csrc/binary_mlp_tc.cu)."""
        from .cost_volume import sw128_tiles

        mlp = self.mlps["s0"]
        W1 = mlp[0].weight.detach().to(dev, torch.float32)
        C = W1.shape[1] - (2 if self.use_prior else 1)
        if C != 64 or W1.shape[0] != 128:
            raise ValueError("the fused binary MLP kernel is built for 64 feature channels and mlp_size 128")

        def split(m):
            hi = m.to(torch.bfloat16)
            return hi, (m - hi.float()).to(torch.bfloat16)

        h1, l1 = split(W1[:, 1:1 + C].contiguous())
        h2, l2 = split(mlp[2].weight.detach().to(dev, torch.float32).contiguous())
        t2h, t2l = sw128_tiles(h2), sw128_tiles(l2)  # two 16 KB K-chunks each
        wimage = torch.cat([sw128_tiles(h1), sw128_tiles(l1), t2h[:16384], t2l[:16384], t2h[16384:],
                            t2l[16384:]]).contiguous()
        vecs = torch.zeros((6, 128), device=dev, dtype=torch.float32)
        vecs[0] = mlp[0].bias.detach()
        vecs[1] = W1[:, 0]
        if self.use_prior:
            vecs[2] = W1[:, 1 + C]
        vecs[3] = mlp[2].bias.detach()
        vecs[4] = mlp[4].weight.detach()[0]
        vecs[5, 0] = mlp[4].bias.detach()[0]
        return wimage, vecs

    def _fused_plan(self, g: Plan, feat: SplitAct):
        import ctypes as C

        class Desc(C.Structure):
            _fields_ = [("feat_hi", C.c_void_p), ("feat_lo", C.c_void_p), ("npix", C.c_longlong), ("HW", C.c_int),
                        ("wimage", C.c_void_p), ("vecs", C.c_void_p), ("use_prior", C.c_int)]

        wimage, vecs = self._pack(g.device)
        d = Desc(feat.hi.data_ptr(), feat.lo.data_ptr(), feat.B * feat.H * feat.W, feat.H * feat.W, wimage.data_ptr(),
                 vecs.data_ptr(), 1 if self.use_prior else 0)
        handle = C.c_void_p()
        _abi.call("b200_binary_mlp_create", C.byref(d), C.byref(handle))
        g._keep += [wimage, vecs, feat, _PlanHandle(handle, "b200_binary_mlp_destroy")]
        return handle

    def plan_val(self, g: Plan, feat: SplitAct, get_depth, num_planes, get_prior=None):
        """BDModel.run_mlp_val for every rendered plane (bd_model.py:293-304, 412-442) as ONE fused tcgen05 kernel:
        the 64-channel feature tile of 128 pixels is loaded once and all planes are evaluated from it.
        get_depth() -> [B, P, H, W] fp32; get_prior() -> [B, 1, H, W] fp32 or None (= -1 everywhere when the model
        uses a prior); returns pred [B, P, H, W] fp32 (logits)."""
        B, H, W, _ = feat.shape
        handle = self._fused_plan(g, feat)
        pred = g.empty((B, num_planes, H, W))

        def op():
            d = get_depth()
            assert tuple(d.shape) == (B, num_planes, H, W) and d.is_contiguous() and d.dtype == torch.float32
            pr = get_prior() if (self.use_prior and get_prior is not None) else None
            if pr is not None:
                assert tuple(pr.shape) == (B, 1, H, W) and pr.is_contiguous() and pr.dtype == torch.float32
            _abi.require_cuda(d, pr)
            _abi.call("b200_binary_mlp_planes", handle, _abi.ptr(d), num_planes, _abi.ptr(pr), _abi.ptr(pred),
                      _abi.stream_ptr())

        g.add(op)
        return pred

    def plan_search(self, g: Plan, feat: SplitAct, get_prior=None, iters=12, min_bound=0.5, max_bound=8.0,
                    first_depth=7.5 / 2.0, get_thresholds=None):
        """The `infer_depth` bisection of BDModel.forward (bd_model.py:273-292) in the same fused kernel: per pixel
        12 evaluations of the MLP at the running query depth, bounds kept in registers.  `get_thresholds()` ->
        (bins, thresholds) fp32 CUDA vectors of equal length (the evaluation's `Thresholder`,
        binary_metrics_utils.py:42-52) or None for the fixed 0.5.  Returns (search_depths, pred of the last
        evaluation), both [B, 1, H, W] fp32."""
        B, H, W, _ = feat.shape
        handle = self._fused_plan(g, feat)
        search = g.empty((B, 1, H, W))
        pred = g.empty((B, 1, H, W))

        def op():
            pr = get_prior() if (self.use_prior and get_prior is not None) else None
            if pr is not None:
                assert tuple(pr.shape) == (B, 1, H, W) and pr.is_contiguous() and pr.dtype == torch.float32
            thr = get_thresholds() if get_thresholds is not None else None
            bins, vals = thr if thr is not None else (None, None)
            if bins is not None:
                assert bins.is_cuda and vals.is_cuda and bins.dtype == vals.dtype == torch.float32
                assert bins.is_contiguous() and vals.is_contiguous() and bins.numel() == vals.numel()
            _abi.call("b200_binary_mlp_search", handle, _abi.ptr(pr), iters, min_bound, max_bound, first_depth,
                      _abi.ptr(bins), _abi.ptr(vals), 0 if bins is None else bins.numel(), _abi.ptr(search),
                      _abi.ptr(pred), _abi.stream_ptr())

        g.add(op)
        return search, pred


class _PlanHandle:
    """Owns a C-side plan object for the life of a launch plan."""

    def __init__(self, handle, destroy):
        self.handle, self.destroy = handle, destroy

    def __del__(self):
        try:
            if self.handle:
                getattr(_abi.load(), self.destroy)(self.handle)
        except Exception:
            pass
