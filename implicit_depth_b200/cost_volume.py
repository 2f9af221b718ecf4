"""B200 drop-ins for the reference's cost/feature-volume managers.

Same constructor arguments, `forward` signature, return tuple, state-dict keys and error
behaviour as `modules/cost_volume.py` (`CostVolumeManager` :17, `FeatureVolumeManager` :369,
`FastFeatureVolumeManager` :718, `EfficientCostVolumeManager` :1149), but the whole
build_cost_volume + argmax runs as hand-written sm_100a kernels behind the C ABI of
`include/b200_planesweep.h`.  Swap in exactly where the reference swaps its fast managers
(`test_bd.py:80-81`):

    model.cost_volume = implicit_depth_b200.to_b200(model.cost_volume)

There is no CPU or PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _abi

FEAT_C = 16
MLP_HID = 128
VIEW_CH = 22
TAIL_CH = 20
MAX_VIEWS = 8


class _PixGrid(nn.Module):
    """Holds `backprojector.pix_coords_13N` so reference checkpoints load (`geometry_utils.py:34-52`)."""

    def __init__(self, height, width):
        super().__init__()
        xx, yy = torch.meshgrid(torch.arange(width), torch.arange(height), indexing="xy")
        pix = torch.stack((xx, yy), 0) + 0.5
        pix = torch.cat([pix, torch.ones_like(pix[:1])], 0).flatten(1).unsqueeze(0)
        self.register_buffer("pix_coords_13N", pix)


class _Eps(nn.Module):
    """Holds `projector.eps` (`geometry_utils.py:74`); the kernels use the same 1e-5 clamp."""

    def __init__(self):
        super().__init__()
        self.register_buffer("eps", torch.tensor(1e-5).view(1, 1, 1))


class MLP(nn.Module):
    """Parameter container with the reference's layout (`modules/networks.py:218-233`):
    `net.0/2/4` are the Linear layers, LeakyReLU(0.01) in between, last layer linear."""

    def __init__(self, channel_list, disable_final_activation=False):
        super().__init__()
        layers = []
        for i in range(len(channel_list) - 1):
            layers.append(nn.Linear(channel_list[i], channel_list[i + 1]))
            layers.append(nn.LeakyReLU(inplace=True))
        if disable_final_activation:
            layers = layers[:-1]
        self.net = nn.Sequential(*layers)


def split_bf16(x):
    """x ~= hi + lo, both bf16 (16 mantissa bits together): the operand format of the tensor-core kernels."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def sw128_tiles(mat_nk):
    """[128, 64*c] bf16 (row = output neuron n, K-major) -> bytes of c consecutive 16 KB tiles in the
    128-byte-swizzled canonical UMMA layout (16-byte chunk j of row r stored at chunk j ^ (r & 7))."""
    n, k = mat_nk.shape
    assert n == 128 and k % 64 == 0 and mat_nk.dtype == torch.bfloat16
    t = mat_nk.view(128, k // 64, 8, 8).permute(1, 0, 2, 3)  # [chunk, row, 16B-chunk, elem]
    r = torch.arange(128, device=mat_nk.device).view(1, 128, 1, 1)
    j = torch.arange(8, device=mat_nk.device).view(1, 1, 8, 1)
    idx = (j ^ (r & 7)).expand(k // 64, 128, 8, 8)
    return torch.gather(t, 2, idx).contiguous().view(torch.uint8).reshape(-1)


def _as_f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _pixel_major(feats, n_img, C, h, w, layout=0):
    """[..., C, h, w] (any batch strides, dense planes) -> one of the kernels' gather layouts (csrc/common.cuh) via
    the layout kernel: 0 = texel records [n_img, h*w, C] (dot-product kernel), 1 = quarter-planar
    [n_img, C/4, h*w, 4] (feature-volume kernels; returned as an opaque buffer of the same shape)."""
    if feats.dtype != torch.float32:
        feats = feats.float()
    f = feats.reshape(n_img, C, h, w) if feats.dim() != 4 else feats
    if f.stride(3) != 1 or f.stride(2) != w:
        f = f.contiguous()
    # images must be evenly strided
    out = torch.empty((n_img, h * w, C), device=f.device, dtype=torch.float32)
    _abi.call("b200_feats_to_pixel_major", _abi.ptr(f), _abi.ptr(out), n_img, C, h * w,
              f.stride(0) if n_img > 1 else C * h * w, f.stride(1), layout, _abi.stream_ptr())
    return out


class B200CostVolumeManager(nn.Module):
    """Dot-product plane-sweep volume; drop-in for `CostVolumeManager` / `EfficientCostVolumeManager`."""

    FEAT_LAYOUT = 0  # gather layout `forward_pixel_major` expects (csrc/common.cuh)

    def __init__(self, matching_height, matching_width, num_depth_bins=64, matching_dim_size=None,
                 num_source_views=None, dot_impl="gather"):
        """dot_impl: "gather" = per-tap global loads through L1 (csrc/cv_dot.cu; the default: 0.32 ms per B=4 launch at
        BASELINE cfg2), "band" = source bands staged in shared memory by TMA (csrc/cv_dot_band.cu; 0.65 ms: with 64-byte
        fp32 texels a 20x20 box is re-used by only 2.5 taps per texel, so the staging traffic exceeds what the hardware L1
        already saves the gather -- measurements in profiles/r02i_volume_kernels.md).  Same results bit for bit."""
        super().__init__()
        if dot_impl not in ("band", "gather"):
            raise ValueError(f"unknown dot_impl {dot_impl!r} (band | gather)")
        self.dot_impl = dot_impl
        self.num_depth_bins = num_depth_bins
        self.matching_height = matching_height
        self.matching_width = matching_width
        self.register_buffer("linear_ramp_1d11", torch.linspace(0, 1, num_depth_bins).view(1, num_depth_bins, 1, 1))
        self.backprojector = _PixGrid(matching_height, matching_width)
        self.projector = _Eps()
        self.max_ctas = 0  # CTA cap of the persistent feature-volume kernel (0 = all SMs); see B200BDModel

    # ---- helpers -----------------------------------------------------------------------
    def _check_shapes(self, cur_feats, src_feats):
        if src_feats.dim() != 5:
            raise ValueError(f"src_feats must be B x K x C x H x W, got {tuple(src_feats.shape)}")
        B, K, C, h, w = src_feats.shape
        if (h, w) != (self.matching_height, self.matching_width):
            raise ValueError(f"feature maps are {h}x{w}, manager was built for "
                             f"{self.matching_height}x{self.matching_width}")
        if tuple(cur_feats.shape) != (B, C, h, w):
            raise ValueError(f"cur_feats shape {tuple(cur_feats.shape)} does not match src_feats {tuple(src_feats.shape)}")
        if C != FEAT_C:
            raise ValueError(f"B200 volume kernels are built for {FEAT_C} feature channels, got {C}")
        if K > MAX_VIEWS:
            raise ValueError(f"at most {MAX_VIEWS} source views, got {K}")
        return B, K, C, h, w

    def _planes_in(self, depth_planes_bdhw, B):
        """User-supplied planes must be constant per (b, d) like the reference's own
        (`cost_volume.py:128-130`); returns a [B, D] tensor or None."""
        if depth_planes_bdhw is None:
            return None
        D = self.num_depth_bins
        if depth_planes_bdhw.shape[0] != B or depth_planes_bdhw.shape[1] != D:
            raise ValueError("depth_planes_bdhw must be B x num_depth_bins x H x W")
        flat = depth_planes_bdhw[:, :, 0, 0]
        if depth_planes_bdhw.stride(2) != 0 or depth_planes_bdhw.stride(3) != 0:
            if not bool((depth_planes_bdhw == flat[:, :, None, None]).all()):
                raise NotImplementedError("per-pixel depth planes are not supported by the B200 kernels")
        return _as_f32c(flat)

    def _prepare(self, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth, planes_in, B, K, D,
                 W1=None, b1=None):
        dev = src_Ks.device
        cams = torch.empty((B, K, 32), device=dev, dtype=torch.float32)
        planes = torch.empty((B, D), device=dev, dtype=torch.float32)
        bias_eff = torch.empty((B, MLP_HID), device=dev, dtype=torch.float32) if W1 is not None else None
        mn = mx = None
        per_frame = 0
        if planes_in is None:
            # the reference broadcasts min/max over the batch (cost_volume.py:117-126): one range or one per frame
            mn, mx = _as_f32c(min_depth.reshape(-1)), _as_f32c(max_depth.reshape(-1))
            if mn.numel() != mx.numel() or mn.numel() not in (1, B):
                raise ValueError(f"min_depth / max_depth must hold 1 or B={B} values, got {mn.numel()} / {mx.numel()}")
            per_frame = 1 if (mn.numel() == B and B > 1) else 0
        _abi.call("b200_volume_prepare", _abi.ptr(_as_f32c(src_Ks)), _abi.ptr(_as_f32c(src_extrinsics)),
                  _abi.ptr(_as_f32c(src_poses)), _abi.ptr(_as_f32c(cur_invK)), _abi.ptr(mn), _abi.ptr(mx),
                  _abi.ptr(planes_in), _abi.ptr(W1), _abi.ptr(b1), _abi.ptr(cams), _abi.ptr(planes),
                  _abi.ptr(bias_eff), B, K, D, FEAT_C, per_frame, _abi.stream_ptr())
        return cams, planes, bias_eff

    # ---- reference API -----------------------------------------------------------------
    def build_cost_volume(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth,
                          max_depth, depth_planes_bdhw=None, return_mask=False):
        cv, _, planes, mask = self.forward(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK,
                                           min_depth, max_depth, depth_planes_bdhw, return_mask)
        return cv, planes, mask

    @torch.no_grad()
    def forward(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                depth_planes_bdhw=None, return_mask=False):
        _abi.require_cuda(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK)
        B, K, C, h, w = self._check_shapes(cur_feats, src_feats)
        cur_pm = _pixel_major(cur_feats, B, C, h, w, self.FEAT_LAYOUT)
        src_pm = _pixel_major(src_feats.reshape(B * K, C, h, w) if src_feats.is_contiguous()
                              else src_feats.contiguous().view(B * K, C, h, w), B * K, C, h, w, self.FEAT_LAYOUT)
        return self.forward_pixel_major(cur_pm, src_pm, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth,
                                        max_depth, depth_planes_bdhw, return_mask, B, K, h, w)

    @torch.no_grad()
    def forward_pixel_major(self, cur_pm, src_pm, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                            depth_planes_bdhw, return_mask, B, K, h, w):
        D = self.num_depth_bins
        N = h * w
        planes_in = self._planes_in(depth_planes_bdhw, B)
        cams, planes, _ = self._prepare(src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth, planes_in,
                                        B, K, D)
        cost = torch.empty((B, D, h, w), device=cur_pm.device, dtype=torch.float32)
        lowest = torch.empty((B, h, w), device=cur_pm.device, dtype=torch.float32)
        _abi.call("b200_cv_dot_band" if self.dot_impl == "band" else "b200_cv_dot", _abi.ptr(cur_pm), _abi.ptr(src_pm),
                  _abi.ptr(cams), _abi.ptr(planes), _abi.ptr(cost), _abi.ptr(lowest), None, B, K, FEAT_C, h, w, D,
                  _abi.stream_ptr())
        planes_bdhw = planes.view(B, D, 1, 1).expand(B, D, h, w) if depth_planes_bdhw is None else depth_planes_bdhw
        return cost, lowest, planes_bdhw, None

    def to_fast(self):
        """The reference swaps in its batched twin here (`cost_volume.py:360-366`); this manager
        already is the fast path."""
        return self


class B200FeatureVolumeManager(B200CostVolumeManager):
    """Metadata-MLP feature volume; drop-in for `FeatureVolumeManager` / `FastFeatureVolumeManager`."""

    FEAT_LAYOUT = 1

    def __init__(self, matching_height, matching_width, num_depth_bins=64, mlp_channels=None, matching_dim_size=16,
                 num_source_views=7, impl="auto"):
        super().__init__(matching_height, matching_width, num_depth_bins)
        K = num_source_views
        cin = matching_dim_size * (1 + K) + (1 + K) + 3 * (1 + K) + K + K + K + 3 * K  # cost_volume.py:405-423
        chans = [cin, 128, 128, 1] if mlp_channels is None else [cin] + list(mlp_channels[1:])
        if chans[1:] != [128, 128, 1]:
            raise ValueError("B200 feature-volume kernels are built for a cin-128-128-1 MLP")
        if matching_dim_size != FEAT_C:
            raise ValueError(f"B200 volume kernels are built for {FEAT_C} feature channels")
        self.num_source_views = K
        self.mlp = MLP(chans, disable_final_activation=True)
        self.impl = impl
        self._packed = None
        self._packed_key = None

    # channel permutation: my view-major order -> reference channel index (cost_volume.py:681-695)
    @staticmethod
    def channel_permutation(K, C=FEAT_C):
        off_depth = C * (K + 1) + K
        off_zd = C * (K + 1) + 2 * K
        off_dot = off_zd + 1
        off_ang = off_dot + K
        off_rays = off_ang + K
        perm = []
        for k in range(K):
            perm += [k * C + c for c in range(C)]
            perm += [off_depth + k, off_dot + k, off_ang + k]
            perm += [off_rays + 3 + 3 * k + i for i in range(3)]
        perm += [K * C + c for c in range(C)]
        perm += [off_rays + i for i in range(3)]
        perm += [off_zd]
        return perm

    @staticmethod
    def tc_channel_layout(K, C=FEAT_C):
        """K-dimension layout of the tensor-core kernel (csrc/fv_tc.cu): reference channel index of every kernel
        channel, -1 for zero padding.  Role A builds views 0..KA-1 + cur[0..7], role B the other views +
        cur[8..15], the current ray and z_d; each role is padded to whole 32-channel halves."""
        import ctypes

        out = (ctypes.c_int * 4)()
        _abi.load().b200_fv_tc_layout(K, out)
        KA, HA, HB, nchunk = out[0], out[1], out[2], out[3]
        ref = B200FeatureVolumeManager.channel_permutation(K, C)  # view-major order: K view blocks, then the tail
        view = lambda k: ref[k * VIEW_CH:(k + 1) * VIEW_CH]
        tail = ref[K * VIEW_CH:]  # cur[0..15], curray[0..2], z_d
        a = [c for k in range(KA) for c in view(k)] + tail[:8]
        b = [c for k in range(KA, K) for c in view(k)] + tail[8:]
        assert len(a) <= 32 * HA and len(b) <= 32 * HB and HA + HB == 2 * nchunk
        a = a + [-1] * (32 * HA - len(a))
        b = b + [-1] * (32 * HB - len(b))
        # the roles' 32-channel halves are interleaved in K (A0 B0 A1 B1 ..., then the longer role's remainder):
        # FvCfg::half_pos in csrc/fv_tc.cu
        halves, hmin = [], min(HA, HB)
        for h in range(hmin):
            halves += [a[32 * h:32 * h + 32], b[32 * h:32 * h + 32]]
        halves += [a[32 * h:32 * h + 32] for h in range(hmin, HA)] + [b[32 * h:32 * h + 32] for h in range(hmin, HB)]
        return [c for half in halves for c in half]

    def _pack(self, device):
        lin = [self.mlp.net[0], self.mlp.net[2], self.mlp.net[4]]
        key = tuple((p.data_ptr(), p._version) for l in lin for p in (l.weight, l.bias)) + (str(device),)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        K = self.num_source_views
        W1 = lin[0].weight.detach().to(device=device, dtype=torch.float32).contiguous()
        perm = torch.tensor(self.channel_permutation(K), device=device, dtype=torch.long)
        kin = VIEW_CH * K + TAIL_CH
        KP = (kin + 15) // 16 * 16
        W1p = torch.zeros((KP, MLP_HID), device=device, dtype=torch.float32)
        W1p[:kin] = W1[:, perm].t()
        layout = torch.tensor(self.tc_channel_layout(K), device=device, dtype=torch.long)
        W1n = torch.zeros((MLP_HID, layout.numel()), device=device, dtype=torch.float32)
        W1n[:, layout >= 0] = W1[:, layout[layout >= 0]]
        W2 = lin[1].weight.detach().to(device=device, dtype=torch.float32)
        h1, l1 = split_bf16(W1n)
        h2, l2 = split_bf16(W2)
        wimage = torch.cat([sw128_tiles(h1), sw128_tiles(l1), sw128_tiles(h2), sw128_tiles(l2)]).contiguous()
        pk = dict(
            wimage=wimage,
            W1=W1, b1=lin[0].bias.detach().to(device=device, dtype=torch.float32).contiguous(),
            W1p=W1p.contiguous(), KP=KP,
            W2t=lin[1].weight.detach().to(device=device, dtype=torch.float32).t().contiguous(),
            b2=lin[1].bias.detach().to(device=device, dtype=torch.float32).contiguous(),
            w3=lin[2].weight.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous(),
            b3=lin[2].bias.detach().to(device=device, dtype=torch.float32).contiguous(),
        )
        self._packed, self._packed_key = pk, key
        return pk

    @torch.no_grad()
    def forward_pixel_major(self, cur_pm, src_pm, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                            depth_planes_bdhw, return_mask, B, K, h, w):
        if K != self.num_source_views:
            raise ValueError(f"manager was built for {self.num_source_views} source views, got {K} "
                             "(the MLP input width depends on it, cost_volume.py:405-423)")
        D = self.num_depth_bins
        N = h * w
        dev = cur_pm.device
        pk = self._pack(dev)
        planes_in = self._planes_in(depth_planes_bdhw, B)
        cams, planes, bias_eff = self._prepare(src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                                               planes_in, B, K, D, pk["W1"], pk["b1"])
        vol = torch.empty((B, D, h, w), device=dev, dtype=torch.float32)
        mask = torch.empty((B, h, w), device=dev, dtype=torch.bool) if return_mask else None
        invK = _as_f32c(cur_invK)
        if self.impl in ("auto", "tc"):
            _abi.call("b200_fv_mlp_tc", _abi.ptr(cur_pm), _abi.ptr(src_pm), _abi.ptr(cams), _abi.ptr(invK),
                      _abi.ptr(planes), _abi.ptr(bias_eff), _abi.ptr(pk["wimage"]), _abi.ptr(pk["b2"]),
                      _abi.ptr(pk["w3"]), _abi.ptr(pk["b3"]), _abi.ptr(vol), _abi.ptr(mask), B, K, FEAT_C, h, w, D,
                      int(self.max_ctas), _abi.stream_ptr())
        elif self.impl == "simt":
            _abi.call("b200_fv_mlp_simt", _abi.ptr(cur_pm), _abi.ptr(src_pm), _abi.ptr(cams), _abi.ptr(invK),
                      _abi.ptr(planes), _abi.ptr(bias_eff), _abi.ptr(pk["W1p"]), _abi.ptr(pk["W2t"]),
                      _abi.ptr(pk["b2"]), _abi.ptr(pk["w3"]), _abi.ptr(pk["b3"]), _abi.ptr(vol), _abi.ptr(mask), B,
                      K, FEAT_C, h, w, D,
                      pk["KP"], _abi.stream_ptr())
        else:
            raise ValueError(f"unknown impl {self.impl!r} (auto | tc | simt)")
        lowest = torch.empty((B, h, w), device=dev, dtype=torch.float32)
        _abi.call("b200_volume_argmax", _abi.ptr(vol), _abi.ptr(planes), _abi.ptr(lowest), None, B, D, N,
                  _abi.stream_ptr())
        planes_bdhw = planes.view(B, D, 1, 1).expand(B, D, h, w) if depth_planes_bdhw is None else depth_planes_bdhw
        return vol, lowest, planes_bdhw, mask


def to_b200(manager):
    """Convert a reference manager (or anything with its attributes) into the B200 drop-in,
    sharing the MLP parameters by reference like the reference's own `to_fast` (`cost_volume.py:708-715`)."""
    if isinstance(manager, B200CostVolumeManager):
        return manager
    h, w, D = manager.matching_height, manager.matching_width, manager.num_depth_bins
    if hasattr(manager, "mlp"):
        first = manager.mlp.net[0]
        K = (first.in_features - 20) // 26
        out = B200FeatureVolumeManager(h, w, num_depth_bins=D, num_source_views=K)
        out.mlp = manager.mlp
    else:
        out = B200CostVolumeManager(h, w, num_depth_bins=D)
    try:
        dev = next(manager.buffers()).device
        out = out.to(dev)
    except StopIteration:
        pass
    return out
