"""CPU ORACLE (test infrastructure, NOT product code) for the plane-sweep volume.

A plain numpy restatement of the reference's cost/feature-volume algorithm.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product (`implicit_depth_b200/`) never does.

Parity status: PINNED.  `tests/golden/gen_golden.py` runs the unmodified reference
classes from `/root/reference` (CostVolumeManager, EfficientCostVolumeManager,
FeatureVolumeManager, FastFeatureVolumeManager) on seeded inputs in the build
container and stores their outputs under `tests/golden/`; `tests/test_oracle_golden.py`
checks every function here against those vectors (the reference itself ships no
tests or known-answer vectors).

All citations are file:line in the reference repository (nianticlabs/implicit-depth).
Every function takes/returns numpy arrays and is dtype-generic: pass float32 inputs
for a like-for-like fp32 restatement or float64 inputs for the arbiter.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- #
# geometry
# --------------------------------------------------------------------------- #
def generate_depth_planes(min_depth, max_depth, num_planes, dtype=np.float32):
    """Log-spaced plane depths, `modules/cost_volume.py:117-126` with the ramp of `:67`.

    z_d = exp(log(zmin) + log(zmax / zmin) * linspace(0, 1, D)[d])
    """
    dt = np.dtype(dtype).type
    ramp = np.linspace(0.0, 1.0, num_planes).astype(dtype)
    lo = np.log(dt(min_depth))
    span = np.log(dt(max_depth) / dt(min_depth))
    return np.exp(lo + span * ramp).astype(dtype)


def pixel_grid(height, width, dtype=np.float32):
    """Homogeneous pixel-centre grid [3, h*w], `utils/geometry_utils.py:34-48` (+0.5 centres)."""
    xx, yy = np.meshgrid(np.arange(width), np.arange(height), indexing="xy")
    pix = np.stack([xx, yy], 0).astype(dtype) + dtype(0.5)
    ones = np.ones((1, height, width), dtype)
    return np.concatenate([pix, ones], 0).reshape(3, -1)


def backproject(depth, invK, pix_3N):
    """`BackprojectDepth.forward`, `utils/geometry_utils.py:54-63`: depth * (invK[:3,:3] @ pix)."""
    cam = invK[:3, :3] @ pix_3N
    return depth * cam  # [3, N]


def project(points_3N, K_44, T_44, eps=1e-5):
    """`Project3D.forward`, `utils/geometry_utils.py:76-89`.

    P = K @ T; c = P[:3] @ [X;1]; z = max(c_z, eps); returns (px, py, z).
    """
    P = (K_44 @ T_44)[:3]
    c = P[:, :3] @ points_3N + P[:, 3:4]
    z = np.maximum(c[2], points_3N.dtype.type(eps))
    return c[0] / z, c[1] / z, z


def pose_distance(pose_44):
    """`pose_distance`, `utils/geometry_utils.py:183-195` (DVMVS pose distance)."""
    dt = pose_44.dtype.type
    R = pose_44[:3, :3]
    t = pose_44[:3, 3]
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    r_meas = np.sqrt(dt(2) * (dt(1) - np.minimum(dt(3), tr) / dt(3)))
    t_meas = np.sqrt((t * t).sum())
    return np.sqrt(t_meas**2 + r_meas**2), r_meas, t_meas


def grid_sample_bilinear_zeros(src_chw, px, py, via_normalised=True):
    """ATen `grid_sampler_2d` (bilinear, zeros padding, align_corners=False) at pixel
    coordinates (px, py), as used at `modules/cost_volume.py:190-198`.

    The reference first normalises `uv = 2 * p * (1/size) - 1` (`:190`) and ATen
    un-normalises `((uv + 1) * size - 1) / 2`; `via_normalised` keeps that exact op
    order (it equals p - 0.5 up to rounding).  Taps outside the image contribute 0;
    weights come from floor().
    """
    C, H, W = src_chw.shape
    dt = src_chw.dtype.type
    if via_normalised:
        u = dt(2) * px * dt(1.0 / W) - dt(1)
        v = dt(2) * py * dt(1.0 / H) - dt(1)
        ix = ((u + dt(1)) * dt(W) - dt(1)) / dt(2)
        iy = ((v + dt(1)) * dt(H) - dt(1)) / dt(2)
    else:
        ix = px - dt(0.5)
        iy = py - dt(0.5)
    x0f = np.floor(ix)
    y0f = np.floor(iy)
    x1f = x0f + dt(1)
    y1f = y0f + dt(1)
    w_nw = (x1f - ix) * (y1f - iy)
    w_ne = (ix - x0f) * (y1f - iy)
    w_sw = (x1f - ix) * (iy - y0f)
    w_se = (ix - x0f) * (iy - y0f)
    big = 2.0**40  # saturate before the int cast (coords reach ~1e8 behind the camera)
    x0 = np.clip(x0f, -big, big).astype(np.int64)
    y0 = np.clip(y0f, -big, big).astype(np.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    flat = src_chw.reshape(C, H * W)
    out = np.zeros((C,) + px.shape, src_chw.dtype)
    for xi, yi, wgt in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        idx = np.where(ok, yi * W + xi, 0)
        vals = flat[:, idx.reshape(-1)].reshape((C,) + px.shape)
        out += vals * np.where(ok, wgt, dt(0))[None]
    return out


# --------------------------------------------------------------------------- #
# dot-product cost volume (simple_cost_volume)
# --------------------------------------------------------------------------- #
def cost_volume_dot(cur_feats, src_feats, src_extrinsics, src_Ks, cur_invK, planes):
    """`CostVolumeManager.build_cost_volume` + `forward`, `modules/cost_volume.py:221-358`.

    cur_feats [B,C,h,w]; src_feats [B,K,C,h,w]; src_extrinsics/src_Ks [B,K,4,4];
    cur_invK [B,4,4]; planes [D].
    Returns cost [B,D,h,w], argmax index [B,h,w] (first max, `:354`), lowest_cost [B,h,w].
    """
    B, K, C, h, w = src_feats.shape
    D = planes.shape[0]
    dt = cur_feats.dtype
    pix = pixel_grid(h, w, dt.type)
    cost = np.zeros((B, D, h, w), dt)
    for b in range(B):
        cur = cur_feats[b].reshape(C, -1)
        for d in range(D):
            X = backproject(planes[d], cur_invK[b], pix)  # :178
            acc = np.zeros(h * w, dt)
            for k in range(K):
                px, py, z = project(X, src_Ks[b, k], src_extrinsics[b, k])  # :182-184
                warped = grid_sample_bilinear_zeros(src_feats[b, k], px, py)  # :192-198
                mask = (z > 0).astype(dt)  # :216 (identically 1, z is clamped to eps)
                acc += (warped * cur).sum(0) * mask  # :302-311
            cost[b, d] = acc.reshape(h, w)
    idx = np.argmax(cost, axis=1)  # :354, first maximal index
    lowest = planes[idx]
    return cost, idx, lowest


# --------------------------------------------------------------------------- #
# MLP feature volume (mlp_feature_volume)
# --------------------------------------------------------------------------- #
def leaky_relu(x, slope):
    return np.where(x >= 0, x, x * x.dtype.type(slope))


def mlp_forward(x, weights, slope=0.01):
    """`MLP`, `modules/networks.py:218-233`: Linear + LeakyReLU(0.01) stack, last layer linear.

    weights = [(W1,b1),(W2,b2),(W3,b3)] with W as stored by nn.Linear ([out,in])."""
    for i, (W, b) in enumerate(weights):
        x = x @ W.T + b
        if i + 1 < len(weights):
            x = leaky_relu(x, slope)
    return x


def _normalize(v3N, eps):
    n = np.sqrt((v3N * v3N).sum(0, keepdims=True))
    return v3N / np.maximum(n, v3N.dtype.type(eps))


def feature_volume_inputs(cur_feats_b, src_feats_b, src_extr_b, src_poses_b, src_Ks_b, cur_invK_b, z_d, pix):
    """The 16*(K+1)+... channel MLP input of one plane of one frame, in the reference's
    exact channel order (`modules/cost_volume.py:681-695`).  Returns ([N, Cin], px, py, z)
    with px/py/z the [K,N] projections (needed for the last-plane mask)."""
    K, C, h, w = src_feats_b.shape
    dt = cur_feats_b.dtype
    N = h * w
    cur = cur_feats_b.reshape(C, N)
    X = backproject(z_d, cur_invK_b, pix)  # :548
    warped, depths, dots, angles, srcrays = [], [], [], [], []
    pxs, pys = [], []
    cur_ray = _normalize(X, 1e-12)  # :618  F.normalize
    for k in range(K):
        px, py, z = project(X, src_Ks_b[k], src_extr_b[k])  # :552-554
        wk = grid_sample_bilinear_zeros(src_feats_b[k], px, py)  # :571-579
        mask = (z > 0).astype(dt)  # :600-601
        t_k = src_poses_b[k, :3, 3:4]
        ray_k = _normalize(X - t_k, 1e-12)  # geometry_utils.py:174-178
        # F.cosine_similarity(eps=1e-5), :657-659 (ATen: divide each by clamped norm, then dot)
        n1 = np.maximum(np.sqrt((cur_ray * cur_ray).sum(0)), dt.type(1e-5))
        n2 = np.maximum(np.sqrt((ray_k * ray_k).sum(0)), dt.type(1e-5))
        ang = ((cur_ray / n1) * (ray_k / n2)).sum(0)
        warped.append(wk)
        depths.append(z)
        dots.append((wk * cur).sum(0) * mask)  # :662-668
        angles.append(ang)
        srcrays.append(ray_k)
        pxs.append(px)
        pys.append(py)
    pd = [pose_distance(src_poses_b[k]) for k in range(K)]  # :505
    ones = np.ones(N, dt)
    chans = []
    for k in range(K):
        chans.extend(warped[k])  # K*C warped visual features
    chans.extend(cur)  # C current features
    chans.extend([ones] * K)  # mask (==1)
    chans.extend(depths)  # clamped z in each source view
    chans.append(np.full(N, z_d, dt))  # plane depth
    chans.extend(dots)
    chans.extend(angles)
    chans.extend(cur_ray)  # 3
    for k in range(K):
        chans.extend(srcrays[k])  # 3 each
    chans.extend([np.full(N, pd[k][0], dt) for k in range(K)])
    chans.extend([np.full(N, pd[k][1], dt) for k in range(K)])
    chans.extend([np.full(N, pd[k][2], dt) for k in range(K)])
    return np.stack(chans, 1).astype(dt), np.stack(pxs), np.stack(pys), np.stack(depths)


def feature_volume_mlp(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, planes, weights,
                       return_mask=True):
    """`FeatureVolumeManager.build_cost_volume` + `forward`, `modules/cost_volume.py:437-706, 324-358`.

    Returns volume [B,D,h,w], argmax index [B,h,w], lowest_cost [B,h,w], overall_mask [B,h,w] bool.
    overall_mask is taken at the LAST plane only (`:603-615`, `:1061-1063`)."""
    B, K, C, h, w = src_feats.shape
    D = planes.shape[0]
    dt = cur_feats.dtype
    pix = pixel_grid(h, w, dt.type)
    vol = np.zeros((B, D, h, w), dt)
    mask_out = np.zeros((B, h, w), bool)
    for b in range(B):
        for d in range(D):
            x, px, py, z = feature_volume_inputs(cur_feats[b], src_feats[b], src_extrinsics[b], src_poses[b],
                                                 src_Ks[b], cur_invK[b], planes[d], pix)
            vol[b, d] = mlp_forward(x, weights)[:, 0].reshape(h, w)
            if d == D - 1 and return_mask:
                inb = (px > 2) & (px < w - 2) & (py > 2) & (py < h - 2)  # get_mask :75-96
                mask_out[b] = ((z > 0).any(0) & inb.any(0)).reshape(h, w)
    idx = np.argmax(vol, axis=1)
    return vol, idx, planes[idx], mask_out


# --------------------------------------------------------------------------- #
# binary occupancy MLP
# --------------------------------------------------------------------------- #
def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def binary_mlp(feature_s0, rendered_depth, weights, prior=None):
    """`BDModel.run_mlp_val` + `BinaryMLPNetwork` s0, `experiment_modules/bd_model.py:412-442`,
    `modules/networks.py:98-104`: per pixel [depth, feat(64) (, prior)] -> 128 -> 128 -> 1, ELU.

    feature_s0 [B,Cf,H,W]; rendered_depth [B,1,H,W]; prior [B,1,H,W] or None."""
    parts = [rendered_depth, feature_s0] + ([prior] if prior is not None else [])
    x = np.concatenate(parts, 1).transpose(0, 2, 3, 1)
    for i, (W, b) in enumerate(weights):
        x = x @ W.T + b
        if i + 1 < len(weights):
            x = elu(x)
    return x.transpose(0, 3, 1, 2)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def thresholder_bins(planes):
    """`Thresholder.__init__`, `utils/binary_metrics_utils.py:42-47`: bin edges halfway between the planes, last 100."""
    bins = np.zeros_like(planes)
    bins[:-1] = (planes[1:] + planes[:-1]) / 2
    bins[-1] = 100.0
    return bins


def binary_search_depth(feature_s0, weights, prior=None, iters=12, min_bound=0.5, max_bound=8.0, thresholder=None):
    """`infer_depth` branch of `BDModel.forward`, `experiment_modules/bd_model.py:273-292`: per-pixel bisection on
    the occupancy logit.  `thresholder` = (bins, thresholds) of the evaluation's depth-dependent `Thresholder`
    (`get_thresholds`: thresholds[torch.bucketize(z, bins)], `binary_metrics_utils.py:50-52`; bucketize with
    right=False = numpy searchsorted side="left") or None for the fixed 0.5.  Returns (search_depths, pred of the
    last evaluation)."""
    B, _, H, W = feature_s0.shape
    dt = feature_s0.dtype
    lo = np.full((B, 1, H, W), min_bound, dt)
    hi = np.full((B, 1, H, W), max_bound, dt)
    z = np.full((B, 1, H, W), 7.5 / 2.0, dt)
    pred = None
    for _ in range(iters):
        pred = binary_mlp(feature_s0, z, weights, prior)
        thr = 0.5 if thresholder is None else thresholder[1][np.searchsorted(thresholder[0], z, side="left")]
        visible = sigmoid(pred) < thr
        hi = np.where(visible, z, hi)
        lo = np.where(visible, lo, z)
        z = (hi + lo) / dt.type(2)
    return z, pred


def sample_prior(rendered_depth, prior_prediction, cam_to_world, prior_world_to_cam, K, invK, return_coords=False):
    """`BDModel.sample_prior`, `experiment_modules/bd_model.py:395-410`: warp the previous prediction into the
    current view through the rendered depth; `F.grid_sample(mode="nearest")` (zeros padding,
    align_corners=False: unnormalise ((g+1)*size-1)/2, round half to even); -1 where rendered depth <= 0
    (the `z > 0` factor of the mask is always true, z being clamped to 1e-5 by Project3D)."""
    B, _, H, W = rendered_depth.shape
    dt = rendered_depth.dtype.type
    out = np.zeros_like(rendered_depth)
    coords = np.zeros((B, 2, H, W), rendered_depth.dtype)  # un-normalised sample positions (arbiter of rounding flips)
    pix = pixel_grid(H, W, rendered_depth.dtype.type)
    for b in range(B):
        cur_to_prior = prior_world_to_cam[b] @ cam_to_world[b]
        X = backproject(rendered_depth[b].reshape(1, -1), invK[b], pix)
        px, py, _ = project(X, K[b], cur_to_prior)
        gx = (px / dt(W) - dt(0.5)) * dt(2)
        gy = (py / dt(H) - dt(0.5)) * dt(2)
        ix = ((gx + dt(1)) * dt(W) - dt(1)) / dt(2)
        iy = ((gy + dt(1)) * dt(H) - dt(1)) / dt(2)
        rx = np.rint(np.clip(ix, -2, W + 1)).astype(np.int64)
        ry = np.rint(np.clip(iy, -2, H + 1)).astype(np.int64)
        ok = (rx >= 0) & (rx < W) & (ry >= 0) & (ry < H)
        flat = prior_prediction[b].reshape(-1)
        v = np.where(ok, flat[np.where(ok, ry * W + rx, 0)], dt(0))
        v = np.where(rendered_depth[b].reshape(-1) > 0, v, dt(-1))
        out[b, 0] = v.reshape(H, W)
        coords[b, 0], coords[b, 1] = ix.reshape(H, W), iy.reshape(H, W)
    return (out, coords) if return_coords else out


# --------------------------------------------------------------------------- #
# input side of the step; pinned by tests/golden/input_side.npz
# --------------------------------------------------------------------------- #
def relative_poses(src_cam_T_world, src_world_T_cam, cur_cam_T_world, cur_world_T_cam):
    """`experiment_modules/bd_model.py:196-204`: src poses [B,K,4,4], current poses [B,4,4] ->
    (src_cam_T_cur_cam, cur_cam_T_src_cam), the managers' `src_extrinsics` / `src_poses`."""
    src_cam_T_cur_cam = src_cam_T_world @ cur_world_T_cam[:, None]   # bd_model.py:200
    cur_cam_T_src_cam = cur_cam_T_world[:, None] @ src_world_T_cam   # bd_model.py:204
    return src_cam_T_cur_cam, cur_cam_T_src_cam


def intrinsics_pyramid(K_s0, levels=5):
    """`datasets/scannet_dataset.py:479-484`: K_s{i} = K with rows 0, 1 divided by 2**i and its
    inverse (`np.linalg.inv`, in the dtype of the input), i < levels.  K_s0 [..., 4, 4] ->
    (K_s, invK_s), each [levels, ..., 4, 4]."""
    Ks, invKs = [], []
    for i in range(levels):
        K_scaled = np.array(K_s0, copy=True)
        K_scaled[..., :2, :] /= 2 ** i
        Ks.append(K_scaled)
        invKs.append(np.linalg.inv(K_scaled))
    return np.stack(Ks), np.stack(invKs)


def scaled_depth_intrinsics(K_file, file_size, depth_size, flip=False, dtype=np.float32):
    """`datasets/scannet_dataset.py:465-477`: intrinsics read from the scan (at `file_size` = (depthWidth,
    depthHeight)), optionally mirrored, rescaled to the configured depth resolution -- the level-0 matrix the
    pyramid starts from."""
    K = np.array(K_file, dtype=dtype)
    if flip:
        K[0, 2] = dtype(file_size[0]) - K[0, 2]
    K[0] *= dtype(depth_size[0] / float(file_size[0]))
    K[1] *= dtype(depth_size[1] / float(file_size[1]))
    return K
