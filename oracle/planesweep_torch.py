"""CPU ORACLE / CPU BASELINE (test infrastructure, NOT product code): the plane-sweep volumes on the
installed torch CPU kernels (multi-threaded grid_sample / matmul), following the reference's per-plane loop
(`modules/cost_volume.py:221-317` and `:437-706`) op for op.  This is the port that `bench.py` times as the
reference arm / `cpu_baseline` (the reference is Python and cannot travel to the GPU box); `tests/` pin it
against the same goldens as the numpy restatement in `oracle/planesweep.py`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _pix(h, w, dtype, device=None):
    xx, yy = torch.meshgrid(torch.arange(w, device=device), torch.arange(h, device=device), indexing="xy")
    p = torch.stack((xx, yy), 0).to(dtype) + 0.5
    return torch.cat([p, torch.ones_like(p[:1])], 0).flatten(1).unsqueeze(0)  # geometry_utils.py:34-48


def depth_planes(min_depth, max_depth, D, dtype=torch.float32):
    ramp = torch.linspace(0, 1, D).to(dtype)
    mn, mx = torch.tensor(min_depth, dtype=dtype), torch.tensor(max_depth, dtype=dtype)
    return torch.exp(torch.log(mn) + torch.log(mx / mn) * ramp)  # cost_volume.py:123-126


def _warp(src_feats, src_extr, src_Ks, cur_invK, z, pix, h, w):
    """cost_volume.py:176-219 for one plane: returns X [B,3,N], pix coords [B*K,2,h,w], depths, warped feats."""
    B, K, C = src_feats.shape[:3]
    X = z * (cur_invK[:, :3, :3] @ pix)  # geometry_utils.py:60-61
    Xh = torch.cat([X, torch.ones_like(X[:, :1])], 1).repeat_interleave(K, 0)
    P = (src_Ks.reshape(-1, 4, 4) @ src_extr.reshape(-1, 4, 4))[:, :3]  # geometry_utils.py:82-84
    c = P @ Xh
    zc = torch.clamp(c[:, 2:], min=1e-5)
    xy = c[:, :2] / zc
    uv = 2 * xy.view(-1, 2, h, w).permute(0, 2, 3, 1) * torch.tensor([1 / w, 1 / h], dtype=c.dtype, device=c.device) - 1
    warped = F.grid_sample(src_feats.reshape(-1, C, h, w), uv, padding_mode="zeros", mode="bilinear",
                           align_corners=False).view(B, K, C, h, w)
    return X, xy.view(B, K, 2, h, w), zc.view(B, K, h, w), warped


def cost_volume_dot(cur_feats, src_feats, src_extr, src_Ks, cur_invK, planes):
    B, K, C, h, w = src_feats.shape
    pix = _pix(h, w, cur_feats.dtype, cur_feats.device)
    out = []
    for z in planes:
        _, _, zc, warped = _warp(src_feats, src_extr, src_Ks, cur_invK, z, pix, h, w)
        out.append(((warped * cur_feats.unsqueeze(1)).sum(2) * (zc > 0).to(warped.dtype)).sum(1, keepdim=True))
    cost = torch.cat(out, 1)
    idx = torch.argmax(cost, 1)
    return cost, idx, planes[idx]


def feature_volume_mlp(cur_feats, src_feats, src_extr, src_poses, src_Ks, cur_invK, planes, weights,
                       return_mask=True):
    """weights: [(W1,b1),(W2,b2),(W3,b3)] torch tensors.  Same outputs as oracle.planesweep.feature_volume_mlp."""
    B, K, C, h, w = src_feats.shape
    dt = cur_feats.dtype
    pix = _pix(h, w, dt, cur_feats.device)
    R = src_poses[..., :3, :3]
    t = src_poses[..., :3, 3]
    tr = R.diagonal(dim1=-1, dim2=-2).sum(-1)
    r_m = torch.sqrt(2 * (1 - torch.minimum(torch.full_like(tr, 3.0), tr) / 3))  # geometry_utils.py:183-195
    t_m = t.norm(dim=-1)
    d_m = torch.sqrt(t_m**2 + r_m**2)
    expand = lambda v: v[:, :, None, None].expand(B, K, h, w)
    vols = []
    mask_out = None
    for di, z in enumerate(planes):
        X, xy, zc, warped = _warp(src_feats, src_extr, src_Ks, cur_invK, z, pix, h, w)
        mask = (zc > 0).to(dt)
        cur_ray = F.normalize(X, dim=1).view(B, 1, 3, h, w)
        src_ray = F.normalize(X.view(B, 1, 3, h * w) - t[..., None], dim=2).view(B, K, 3, h, w)
        ang = F.cosine_similarity(cur_ray.expand(B, K, 3, h, w), src_ray, dim=2, eps=1e-5)
        dot = (warped * cur_feats.unsqueeze(1)).sum(2) * mask
        feats = torch.cat([warped.flatten(1, 2), cur_feats, mask, zc, torch.full((B, 1, h, w), float(z), dtype=dt, device=cur_feats.device), dot,
                           ang, cur_ray.flatten(1, 2), src_ray.flatten(1, 2), expand(d_m), expand(r_m), expand(t_m)],
                          1)  # cost_volume.py:681-695
        x = feats.permute(0, 2, 3, 1)
        for i, (W, b) in enumerate(weights):
            x = F.linear(x, W, b)
            if i + 1 < len(weights):
                x = F.leaky_relu(x, 0.01)
        vols.append(x.squeeze(-1).unsqueeze(1))
        if return_mask and di == len(planes) - 1:
            inb = (xy[:, :, 0] > 2) & (xy[:, :, 0] < w - 2) & (xy[:, :, 1] > 2) & (xy[:, :, 1] < h - 2)
            mask_out = (zc > 0).any(1) & inb.any(1)
    vol = torch.cat(vols, 1)
    idx = torch.argmax(vol, 1)
    return vol, idx, planes[idx], mask_out


def mask_edge_distance(src_extr, src_Ks, cur_invK, z_last, h, w):
    """Arbiter for `overall_mask_bhw` (`get_mask`, cost_volume.py:75-96, taken at the last plane :603-615): per pixel
    the distance, in pixels of the source image, from the nearest view's projection to the nearest edge of the open
    window 2 < px < w-2, 2 < py < h-2 -- a pixel whose mask bit differs from the reference's is a rounding flip only
    if this distance is ~0.  Pass fp64 tensors.  Returns [B,h,w]."""
    B, K = src_extr.shape[:2]
    pix = _pix(h, w, cur_invK.dtype, cur_invK.device)
    X = z_last * (cur_invK[:, :3, :3] @ pix)
    Xh = torch.cat([X, torch.ones_like(X[:, :1])], 1).repeat_interleave(K, 0)
    P = (src_Ks.reshape(-1, 4, 4) @ src_extr.reshape(-1, 4, 4))[:, :3]
    c = P @ Xh
    xy = (c[:, :2] / torch.clamp(c[:, 2:], min=1e-5)).view(B, K, 2, h, w)
    px, py = xy[:, :, 0], xy[:, :, 1]
    d = torch.stack([(px - 2).abs(), (px - (w - 2)).abs(), (py - 2).abs(), (py - (h - 2)).abs()], 0).amin(0)
    return d.amin(1)
