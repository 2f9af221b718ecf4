"""Input side of the step: one staging buffer per batch instead of per-tensor `.cuda().float()`.

The reference moves a batch to the GPU tensor by tensor (`to_gpu`, utils/generic_utils.py:139-146: ~30 blocking copies
per frame tuple, most of them 64-byte matrices the forward never reads), multiplies the poses with two batched
matmuls (experiment_modules/bd_model.py:196-204) and concatenates current and source images for the matching encoder
(bd_model.py:162-165).  Here a batch is laid out ONCE, on the host, in the order the kernels read it:

    images [B*(K+1), 3, H, W]   current frames first, then the source frames -- already the matching encoder's batch
    rendered_depth [B, P, H/2, W/2]
    src_K [B,K,4,4] | cur_invK [B,4,4] | src_cam_T_world, src_world_T_cam [B,K,4,4] | cur_cam_T_world,
    cur_world_T_cam [B,4,4]  (+ temporal: prior_prediction [B,1,H/2,W/2], prior_cam_T_world, K_s0, invK_s0 [B,4,4])

`FrameStaging.host_frame()` hands the data loader dictionaries of views into one pinned buffer (the reference's keys, so
a collate function fills them in place); `upload()` is ONE `cudaMemcpyAsync`; the device-side dictionaries are views
into the device copy, `B200BDModel.forward` recognises them (`StagedDict`) and lets its CUDA graph read the slot
directly: no per-tensor copies, no `torch.cat`, and the two pose products are one `b200_relative_poses` launch.
"""
from __future__ import annotations

import torch

ALIGN = 64  # floats: every field starts on a 256-byte boundary (TMA / 16-byte vector loads need 16)


class StagedDict(dict):
    """A batch dictionary whose tensors are views into one staging buffer (`.frame` names it)."""

    frame = None


class StagedFrame:
    """One staged batch: `buf` (flat fp32, pinned host or device) plus `cur` / `src` dictionaries of views."""

    def __init__(self, staging, buf):
        self.staging, self.buf = staging, buf
        f = {name: buf[off:off + n].view(shape) for name, (off, n, shape) in staging.fields.items()}
        self.fields = f
        B, K, ms = staging.B, staging.K, staging.matching_scale
        self.cur, self.src = StagedDict(), StagedDict()
        self.cur.frame = self.src.frame = self
        self.cur["image_b3hw"] = f["images"][:B]
        self.src["image_b3hw"] = f["images"][B:].view(B, K, 3, staging.H, staging.W)
        self.cur["rendered_depth"] = f["rendered_depth"]
        self.src[f"K_s{ms}_b44"] = f["src_K"]
        self.cur[f"invK_s{ms}_b44"] = f["cur_invK"]
        self.src["cam_T_world_b44"] = f["src_cam_T_world"]
        self.src["world_T_cam_b44"] = f["src_world_T_cam"]
        self.cur["cam_T_world_b44"] = f["cur_cam_T_world"]
        self.cur["world_T_cam_b44"] = f["cur_world_T_cam"]
        if staging.temporal:
            self.cur["prior_prediction"] = f["prior_prediction"]
            self.cur["prior_cam_T_world"] = f["prior_cam_T_world"]
            self.cur["K_s0_b44"] = f["K_s0"]
            self.cur["invK_s0_b44"] = f["invK_s0"]

    def fill(self, cur_data, src_data):
        """Collate arbitrary (numpy / torch, any float dtype) batch dictionaries into the staged views; keys the
        forward does not read are ignored.  Returns self."""
        for dst, srcd in ((self.cur, cur_data), (self.src, src_data)):
            for k, view in dst.items():
                if k not in srcd or srcd[k] is None:
                    raise KeyError(f"batch dictionary lacks {k!r}, which the forward reads")
                v = srcd[k]
                v = v if isinstance(v, torch.Tensor) else torch.as_tensor(v)
                if tuple(v.shape) != tuple(view.shape):
                    raise ValueError(f"{k}: staged shape {tuple(view.shape)}, got {tuple(v.shape)}")
                view.copy_(v)  # casts to fp32 like to_gpu's .float()
        return self


class FrameStaging:
    """Layout of one staged batch for the signature (B, K, H, W, P)."""

    def __init__(self, B, K, H, W, P=8, temporal=False, matching_scale=1):
        self.B, self.K, self.H, self.W, self.P = B, K, H, W, P
        self.temporal, self.matching_scale = bool(temporal), matching_scale
        spec = [
            ("images", (B * (K + 1), 3, H, W)),
            ("rendered_depth", (B, P, H // 2, W // 2)),
            ("src_K", (B, K, 4, 4)),
            ("cur_invK", (B, 4, 4)),
            ("src_cam_T_world", (B, K, 4, 4)),
            ("src_world_T_cam", (B, K, 4, 4)),
            ("cur_cam_T_world", (B, 4, 4)),
            ("cur_world_T_cam", (B, 4, 4)),
        ]
        if self.temporal:
            spec += [("prior_prediction", (B, 1, H // 2, W // 2)), ("prior_cam_T_world", (B, 4, 4)),
                     ("K_s0", (B, 4, 4)), ("invK_s0", (B, 4, 4))]
        self.fields, off = {}, 0
        for name, shape in spec:
            n = 1
            for s in shape:
                n *= s
            self.fields[name] = (off, n, shape)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.nbytes = off * 4

    @classmethod
    def for_batch(cls, cur_data, src_data, matching_scale=1):
        B, K = src_data["image_b3hw"].shape[:2]
        H, W = cur_data["image_b3hw"].shape[-2:]
        return cls(B, K, H, W, cur_data["rendered_depth"].shape[1],
                   temporal=cur_data.get("prior_prediction", None) is not None, matching_scale=matching_scale)

    def host_frame(self, pin=True):
        """A host-side frame; `pin=False` only for CPU-only unit tests (no CUDA context to pin with)."""
        buf = torch.zeros(self.numel, dtype=torch.float32)
        return StagedFrame(self, buf.pin_memory() if pin else buf)

    def device_frame(self, device):
        """A device-side slot.  Allocate a few and REUSE them (`pipeline.FramePipeline` keeps two): the forward
        captures one CUDA graph per slot address, so a fresh slot per batch would re-capture every call."""
        # torch.empty: no fill kernel on the caller's stream that could race with an upload issued on a copy stream
        # (the alignment gaps between the fields are never read)
        return StagedFrame(self, torch.empty(self.numel, dtype=torch.float32, device=device))

    @staticmethod
    def upload(host_frame, device_frame, stream=None):
        """One asynchronous H2D copy of the whole batch on `stream` (default: current).  Returns the byte count."""
        if host_frame.staging.fields != device_frame.staging.fields:
            raise ValueError("host and device frames have different layouts")
        if stream is None:
            device_frame.buf.copy_(host_frame.buf, non_blocking=True)
        else:
            with torch.cuda.stream(stream):
                device_frame.buf.copy_(host_frame.buf, non_blocking=True)
        return host_frame.staging.nbytes


def relative_poses(src_cam_T_world, src_world_T_cam, cur_cam_T_world, cur_world_T_cam):
    """(src_cam_T_cur_cam, cur_cam_T_src_cam) of bd_model.py:196-204 in one launch; fp32 CUDA tensors."""
    from . import _abi

    _abi.require_cuda(src_cam_T_world, src_world_T_cam, cur_cam_T_world, cur_world_T_cam)
    f = lambda t: (t if t.dtype == torch.float32 else t.float()).contiguous()
    a, b, c, d = f(src_cam_T_world), f(src_world_T_cam), f(cur_cam_T_world), f(cur_world_T_cam)
    B, K = a.shape[:2]
    if tuple(b.shape) != (B, K, 4, 4) or tuple(c.shape) != (B, 4, 4) or tuple(d.shape) != (B, 4, 4) \
            or tuple(a.shape) != (B, K, 4, 4):
        raise ValueError("relative_poses expects src poses [B,K,4,4] and current poses [B,4,4]")
    s2c, c2s = torch.empty_like(a), torch.empty_like(a)
    _abi.call("b200_relative_poses", _abi.ptr(a), _abi.ptr(b), _abi.ptr(c), _abi.ptr(d), _abi.ptr(s2c),
              _abi.ptr(c2s), B, K, _abi.stream_ptr())
    return s2c, c2s


def intrinsics_pyramid(K_s0, levels=5):
    """`K_s{i}_b44`, `invK_s{i}_b44` for i < levels from the level-0 intrinsics (scannet_dataset.py:479-484) in one
    launch.  K_s0 [..., 4, 4] fp32 CUDA; returns two lists of tensors of K_s0's shape."""
    from . import _abi

    _abi.require_cuda(K_s0)
    if K_s0.shape[-2:] != (4, 4):
        raise ValueError("intrinsics_pyramid expects [..., 4, 4] matrices")
    k0 = (K_s0 if K_s0.dtype == torch.float32 else K_s0.float()).contiguous()
    n = k0.numel() // 16
    Ks = torch.empty((levels,) + tuple(k0.shape), device=k0.device, dtype=torch.float32)
    invKs = torch.empty_like(Ks)
    _abi.call("b200_intrinsics_pyramid", _abi.ptr(k0), _abi.ptr(Ks), _abi.ptr(invKs), n, levels, _abi.stream_ptr())
    return list(Ks.unbind(0)), list(invKs.unbind(0))
