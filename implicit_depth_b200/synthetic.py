"""Seeded synthetic inputs with the schema the reference's forward reads.

Mirrors the dictionaries built by `datasets/generic_mvs_dataset.py:742-807` and the
intrinsics pyramid of `datasets/scannet_dataset.py:476-486`.  numpy's
PCG64 generator is used so the very same tensors are produced in the build container
(golden generation) and on the GPU box (parity tests, bench).
"""
from __future__ import annotations

import numpy as np


def _rot(axis, angle):
    axis = axis / np.linalg.norm(axis)
    x, y, z = axis
    c, s = np.cos(angle), np.sin(angle)
    C = 1 - c
    return np.array([[c + x * x * C, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, c + y * y * C, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, c + z * z * C]])


def pose_distance_np(pose):
    """`utils/geometry_utils.py:183-195`."""
    R, t = pose[:3, :3], pose[:3, 3]
    r = np.sqrt(2 * (1 - min(3.0, np.trace(R)) / 3))
    tm = np.linalg.norm(t)
    return np.sqrt(tm * tm + r * r)


def make_intrinsics(image_h, image_w, num_scales=5):
    """K_s{i} / invK_s{i} pyramid: K_s0 is for (W/2, H/2), K_si[:2] = K_s0[:2] / 2^i."""
    K0 = np.eye(4)
    K0[0, 0] = K0[1, 1] = 0.9 * (image_w / 2)
    K0[0, 2] = image_w / 4
    K0[1, 2] = image_h / 4
    Ks, invKs = [], []
    for i in range(num_scales):
        K = K0.copy()
        K[:2] /= 2**i
        Ks.append(K)
        invKs.append(np.linalg.inv(K))
    return Ks, invKs


def make_poses(rng, num_src, max_t=0.3, max_tz=0.05, max_rot_deg=10.0):
    """A reference world_T_cam plus `num_src` DVMVS-like neighbours, sorted by pose distance
    (`datasets/generic_mvs_dataset.py:791-807`)."""
    world_T_cur = np.eye(4)
    world_T_cur[:3, :3] = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
    world_T_cur[:3, 3] = rng.uniform(-1, 1, size=3)
    rel = []
    for _ in range(num_src):
        T = np.eye(4)
        T[:3, :3] = _rot(rng.normal(size=3), np.deg2rad(rng.uniform(0, max_rot_deg)))
        T[:3, 3] = [rng.uniform(-max_t, max_t), rng.uniform(-max_t, max_t), rng.uniform(-max_tz, max_tz)]
        rel.append(T)  # cur_T_src
    rel.sort(key=pose_distance_np)
    world_T_src = [world_T_cur @ T for T in rel]
    return world_T_cur, world_T_src


def make_volume_inputs(seed, B, K, C, h, w, image_scale=4, dtype=np.float32):
    """Inputs of the cost-volume managers' `forward` (`modules/cost_volume.py:324-336`):
    instance-normalised random features + camera matrices at matching resolution."""
    rng = np.random.default_rng(seed)
    H, W = h * image_scale, w * image_scale
    Ks, invKs = make_intrinsics(H, W)
    cur = rng.standard_normal((B, C, h, w))
    src = rng.standard_normal((B, K, C, h, w))
    # smooth a little so neighbouring texels correlate like CNN features, then instance-normalise
    for t in (cur, src):
        t += 0.5 * np.roll(t, 1, -1) + 0.5 * np.roll(t, 1, -2)
        t -= t.mean((-1, -2), keepdims=True)
        t /= t.std((-1, -2), keepdims=True)
    extr = np.zeros((B, K, 4, 4))
    poses = np.zeros((B, K, 4, 4))
    for b in range(B):
        world_T_cur, world_T_src = make_poses(rng, K)
        for k in range(K):
            poses[b, k] = np.linalg.inv(world_T_cur) @ world_T_src[k]  # cur_T_src
            extr[b, k] = np.linalg.inv(world_T_src[k]) @ world_T_cur  # src_T_cur
    srcK = np.broadcast_to(Ks[1], (B, K, 4, 4)).copy()
    invK = np.broadcast_to(invKs[1], (B, 4, 4)).copy()
    out = dict(cur_feats=cur, src_feats=src, src_extrinsics=extr, src_poses=poses, src_Ks=srcK, cur_invK=invK)
    return {k: np.ascontiguousarray(v.astype(dtype)) for k, v in out.items()}


def make_frame_batch(seed, B, K, image_h, image_w, num_rendered=8, temporal=False, dtype=np.float32):
    """(cur_data, src_data) numpy dictionaries with the keys `BDModel.forward` reads
    (`experiment_modules/bd_model.py:186-194, 296, 423-430`)."""
    rng = np.random.default_rng(seed)
    Ks, invKs = make_intrinsics(image_h, image_w)
    cur = {"image_b3hw": rng.standard_normal((B, 3, image_h, image_w))}
    src = {"image_b3hw": rng.standard_normal((B, K, 3, image_h, image_w))}
    for i in range(5):
        cur[f"K_s{i}_b44"] = np.broadcast_to(Ks[i], (B, 4, 4)).copy()
        cur[f"invK_s{i}_b44"] = np.broadcast_to(invKs[i], (B, 4, 4)).copy()
        src[f"K_s{i}_b44"] = np.broadcast_to(Ks[i], (B, K, 4, 4)).copy()
        src[f"invK_s{i}_b44"] = np.broadcast_to(invKs[i], (B, K, 4, 4)).copy()
    wTc = np.zeros((B, 4, 4))
    wTs = np.zeros((B, K, 4, 4))
    for b in range(B):
        a, s = make_poses(rng, K)
        wTc[b] = a
        wTs[b] = np.stack(s)
    cur["world_T_cam_b44"] = wTc
    cur["cam_T_world_b44"] = np.linalg.inv(wTc)
    src["world_T_cam_b44"] = wTs
    src["cam_T_world_b44"] = np.linalg.inv(wTs)
    planes = np.linspace(1.5, 5.0, num_rendered)
    cur["rendered_depth"] = np.broadcast_to(planes[None, :, None, None],
                                            (B, num_rendered, image_h // 2, image_w // 2)).copy()
    cur["depth_b1hw"] = rng.uniform(0.5, 3.5, size=(B, 1, image_h // 2, image_w // 2))
    if temporal:
        prior = rng.uniform(0, 1, size=(B, 1, image_h // 2, image_w // 2))
        prior[..., :4, :] = prior[..., -4:, :] = -1
        prior[..., :, :4] = prior[..., :, -4:] = -1
        cur["prior_prediction"] = prior
        T = np.eye(4)
        T[:3, :3] = _rot(rng.normal(size=3), np.deg2rad(3.0))
        T[:3, 3] = rng.uniform(-0.05, 0.05, size=3)
        cur["prior_cam_T_world"] = np.linalg.inv(wTc @ T[None])
    cast = lambda d: {k: np.ascontiguousarray(v.astype(dtype)) for k, v in d.items()}
    return cast(cur), cast(src)


def init_model_weights(model, seed=0):
    """Seeded 'random weights' (BASELINE.json configs): torch default init under `seed`, plus non-trivial
    BatchNorm statistics so that BN folding is exercised.  Returns a float64 checksum of all tensors."""
    import torch

    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() > 1:
                fan_in = p[0].numel()
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * (3.0 / fan_in) ** 0.5)
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.1)
        for name, m in model.named_modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
    return float(sum(v.double().abs().sum() for v in model.state_dict().values() if v.dtype.is_floating_point))


def make_net_inputs(seed, B=1, image_h=96, image_w=128, D=16, enc_ch=(24, 48, 64, 160, 256)):
    """Random image-encoder pyramid + cost volume for the conv-network tests."""
    rng = np.random.default_rng(seed)
    enc = [rng.standard_normal((B, c, image_h // 2 ** (i + 1), image_w // 2 ** (i + 1))).astype(np.float32)
           for i, c in enumerate(enc_ch)]
    cv = rng.standard_normal((B, D, image_h // 4, image_w // 4)).astype(np.float32)
    img = rng.standard_normal((B, 3, image_h, image_w)).astype(np.float32)
    return enc, cv, img
