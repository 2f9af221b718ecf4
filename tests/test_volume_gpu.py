"""GPU parity tests of the plane-sweep volume kernels, called through the reference-shaped
managers (which go through the C ABI).  Tolerance: 1e-3 relative (max|diff| / max|ref|) as
BASELINE.json's north_star states; the plane index argmax must be exact except at reference
near-ties of the stored fp64 arbiter (the two candidate planes differ by < 1e-5 of the volume's range in the
reference's own fp64 run); every case records n_bad / n_near / margins (profiles/r02_parity_exactness.jsonl)."""
import numpy as np
import pytest
import torch

from implicit_depth_b200 import B200CostVolumeManager, B200FeatureVolumeManager, synthetic
from oracle import planesweep as O

from cases import VOLUME_CASES, argmax_exactness, load_golden, mask_exactness, mlp_weights, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def dev(inp):
    return {k: torch.from_numpy(v).cuda() for k, v in inp.items()}


def depth_range():
    return (torch.tensor(0.25, device="cuda").view(1, 1, 1, 1), torch.tensor(5.0, device="cuda").view(1, 1, 1, 1))


def planes_to_idx(lowest, planes):
    return np.abs(lowest[..., None] - planes[None, None, None, :]).argmin(-1)


def fp64_arbiter(g, key, inp, h, weights=None):
    """The reference's own fp64 volume: stored whole for the small cases; the cfg2 fixture keeps a ::4 sample (file
    size), so there the torch port -- pinned against that sample in tests/test_oracle_golden.py -- recomputes the full
    map in fp64 on the host."""
    if g[key].shape[2] == h:
        return g[key]
    from oracle import planesweep_torch as PT

    t = {k: torch.from_numpy(v).double() for k, v in inp.items()}
    planes = PT.depth_planes(0.25, 5.0, g["planes"].shape[0], torch.float64)
    if weights is None:
        vol = PT.cost_volume_dot(t["cur_feats"], t["src_feats"], t["src_extrinsics"], t["src_Ks"], t["cur_invK"],
                                 planes)[0]
    else:
        W = [(torch.from_numpy(a).double(), torch.from_numpy(b).double()) for a, b in weights]
        vol = PT.feature_volume_mlp(t["cur_feats"], t["src_feats"], t["src_extrinsics"], t["src_poses"], t["src_Ks"],
                                    t["cur_invK"], planes, W, False)[0]
    vol = vol.numpy()
    st = h // g[key].shape[2]
    assert np.abs(vol[:, :, ::st, ::st] - g[key]).max() < 1e-9 * np.abs(g[key]).max()
    return vol


def load_mlp(mgr, g):
    with torch.no_grad():
        for i, li in enumerate((0, 2, 4)):
            mgr.mlp.net[li].weight.copy_(torch.from_numpy(g[f"mlp_w{i}"]))
            mgr.mlp.net[li].bias.copy_(torch.from_numpy(g[f"mlp_b{i}"]))


@pytest.mark.parametrize("dot_impl", ["band", "gather"])
@pytest.mark.parametrize("name", list(VOLUME_CASES))
def test_dot_volume_vs_reference_golden(name, dot_impl):
    seed, B, K, C, h, w, D = VOLUME_CASES[name]
    g = load_golden(name)
    inp = synthetic.make_volume_inputs(seed, B, K, C, h, w)
    t = dev(inp)
    mgr = B200CostVolumeManager(h, w, num_depth_bins=D, dot_impl=dot_impl).cuda()
    mn, mx = depth_range()
    cost, lowest, planes_bdhw, mask = mgr(min_depth=mn, max_depth=mx, **t)
    assert mask is None and tuple(cost.shape) == (B, D, h, w) and tuple(planes_bdhw.shape) == (B, D, h, w)
    cost, lowest = cost.cpu().numpy(), lowest.cpu().numpy()
    planes = planes_bdhw[0, :, 0, 0].cpu().numpy()
    np.testing.assert_allclose(planes, g["planes"], rtol=1e-6)
    assert rel_err(cost, g["dot_cost"]) < TOL
    idx = planes_to_idx(lowest, planes)
    np.testing.assert_array_equal(idx, np.argmax(cost, 1))  # kernel argmax == first max of its own volume
    n_bad, n_near = argmax_exactness(f"dot_volume[{dot_impl}]/{name}", idx, np.argmax(g["dot_cost"], 1),
                                     fp64_arbiter(g, "dot_cost_f64", inp, h))
    assert n_bad == n_near, f"{n_bad} argmax mismatches, only {n_near} at fp64 near-ties"


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("name", list(VOLUME_CASES))
def test_feature_volume_vs_reference_golden(name, impl):
    seed, B, K, C, h, w, D = VOLUME_CASES[name]
    g = load_golden(name)
    inp = synthetic.make_volume_inputs(seed, B, K, C, h, w)
    t = dev(inp)
    mgr = B200FeatureVolumeManager(h, w, num_depth_bins=D, num_source_views=K, impl=impl).cuda()
    load_mlp(mgr, g)
    mn, mx = depth_range()
    vol, lowest, planes_bdhw, mask = mgr(min_depth=mn, max_depth=mx, return_mask=True, **t)
    assert mask.dtype == torch.bool
    vol, lowest, mask = vol.cpu().numpy(), lowest.cpu().numpy(), mask.cpu().numpy()
    assert rel_err(vol, g["fv_vol"]) < TOL
    from oracle import planesweep_torch as PT

    d64 = {k: torch.from_numpy(v).double() for k, v in inp.items()}
    edge = PT.mask_edge_distance(d64["src_extrinsics"], d64["src_Ks"], d64["cur_invK"],
                                 float(PT.depth_planes(0.25, 5.0, D, torch.float64)[-1]), h, w).numpy()
    m_bad, m_edge = mask_exactness(f"feature_volume[{impl}]/{name}", mask, g["fv_mask"], edge)
    assert m_bad == m_edge, f"{m_bad} mask flips, only {m_edge} on the window edge"  # |px - edge| < 1e-3 px
    planes = planes_bdhw[0, :, 0, 0].cpu().numpy()
    idx = planes_to_idx(lowest, planes)
    np.testing.assert_array_equal(idx, np.argmax(vol, 1))
    n_bad, n_near = argmax_exactness(f"feature_volume[{impl}]/{name}", idx, np.argmax(g["fv_vol"], 1),
                                     fp64_arbiter(g, "fv_vol_f64", inp, h, mlp_weights(g)))
    assert n_bad == n_near, f"{n_bad} argmax mismatches, only {n_near} at fp64 near-ties"


@pytest.mark.parametrize("impl", ["tc", "simt"])
@pytest.mark.parametrize("shape", [(1, 1, 7, 9, 1), (2, 4, 33, 17, 3), (1, 8, 5, 130, 2), (3, 5, 11, 40, 4),
                                   (1, 6, 16, 16, 9)])
def test_ragged_shapes_vs_oracle(shape, impl):
    """Edge sizes: one view / one plane, N not a multiple of any tile, maximum view count."""
    B, K, h, w, D = shape
    inp = synthetic.make_volume_inputs(77 + K, B, K, 16, h, w)
    t = dev(inp)
    mn, mx = depth_range()
    mgr = B200CostVolumeManager(h, w, num_depth_bins=D).cuda()
    cost, lowest, planes_bdhw, _ = mgr(min_depth=mn, max_depth=mx, **t)
    planes = planes_bdhw[0, :, 0, 0].cpu().numpy()
    ref, ridx, rlow = O.cost_volume_dot(inp["cur_feats"], inp["src_feats"], inp["src_extrinsics"], inp["src_Ks"],
                                        inp["cur_invK"], planes)
    assert rel_err(cost.cpu().numpy(), ref) < TOL
    fv = B200FeatureVolumeManager(h, w, num_depth_bins=D, num_source_views=K, impl=impl).cuda()
    torch.manual_seed(3)
    for p in fv.parameters():
        torch.nn.init.normal_(p, std=0.1)
    W = [(fv.mlp.net[i].weight.detach().cpu().numpy(), fv.mlp.net[i].bias.detach().cpu().numpy()) for i in (0, 2, 4)]
    vol, lowest, _, mask = fv(min_depth=mn, max_depth=mx, return_mask=True, **t)
    rvol, _, _, rmask = O.feature_volume_mlp(inp["cur_feats"], inp["src_feats"], inp["src_extrinsics"],
                                             inp["src_poses"], inp["src_Ks"], inp["cur_invK"], planes, W)
    assert rel_err(vol.cpu().numpy(), rvol) < TOL
    assert (mask.cpu().numpy() != rmask).mean() < 5e-3


def test_noncontiguous_cur_feats_and_user_planes():
    """cur_feats arrives as the strided slice matching_feats[:, 0] (bd_model.py:170); planes may be
    passed explicitly as an expanded view (cost_volume.py:128-130)."""
    seed, B, K, C, h, w, D = VOLUME_CASES["small_24x32_k7_d8"]
    g = load_golden("small_24x32_k7_d8")
    inp = synthetic.make_volume_inputs(seed, B, K, C, h, w)
    t = dev(inp)
    allf = torch.cat([t["cur_feats"][:, None], t["src_feats"]], 1)
    t["cur_feats"] = allf[:, 0]
    assert not t["cur_feats"].is_contiguous()
    planes_bdhw = torch.from_numpy(g["planes"]).cuda().view(1, D, 1, 1).expand(B, D, h, w)
    mgr = B200CostVolumeManager(h, w, num_depth_bins=D).cuda()
    mn, mx = depth_range()
    cost, lowest, pl, _ = mgr(min_depth=mn, max_depth=mx, depth_planes_bdhw=planes_bdhw, **t)
    assert rel_err(cost.cpu().numpy(), g["dot_cost"]) < TOL
    assert pl is planes_bdhw
    # with the reference's own plane values the gathered depth must be one of them, bit for bit
    assert np.isin(lowest.cpu().numpy(), g["planes"]).all()


def test_all_views_out_of_frustum_gives_zero_and_first_plane():
    seed, B, K, C, h, w, D = VOLUME_CASES["ragged_20x36_k3_d5"]
    inp = synthetic.make_volume_inputs(seed, B, K, C, h, w)
    flip = np.diag([-1.0, 1.0, -1.0, 1.0]).astype(np.float32)
    inp["src_extrinsics"] = flip @ inp["src_extrinsics"]
    mgr = B200CostVolumeManager(h, w, num_depth_bins=D).cuda()
    mn, mx = depth_range()
    cost, lowest, pl, _ = mgr(min_depth=mn, max_depth=mx, **dev(inp))
    assert float(cost.abs().max()) == 0.0
    assert torch.equal(lowest, pl[:, 0])


def test_linearity_and_batch_invariance_full_size():
    """Size-independent properties at the BASELINE cfg2 shape (B=4, 96x128, K=7, D=64): the dot
    volume is linear in the current features, and a frame's result does not depend on what else is
    in the batch (the reference runs its encoder unbatched for exactly this reason,
    depth_model.py:235-241)."""
    B, K, C, h, w, D = 4, 7, 16, 96, 128, 64
    t = dev(synthetic.make_volume_inputs(2000, B, K, C, h, w))
    mgr = B200CostVolumeManager(h, w, num_depth_bins=D).cuda()
    mn, mx = depth_range()
    c1, l1, _, _ = mgr(min_depth=mn, max_depth=mx, **t)
    t2 = dict(t)
    t2["cur_feats"] = t["cur_feats"] * 2.0
    c2, l2, _, _ = mgr(min_depth=mn, max_depth=mx, **t2)
    assert torch.equal(c2, 2.0 * c1) and torch.equal(l1, l2)  # scaling by 2 is exact in fp32
    one = {k: v[1:2].contiguous() for k, v in t.items()}
    c3, l3, _, _ = mgr(min_depth=mn, max_depth=mx, **one)
    assert torch.equal(c3, c1[1:2]) and torch.equal(l3, l1[1:2])
    torch.manual_seed(11)
    ref_mlp = B200FeatureVolumeManager(h, w, num_depth_bins=D).mlp
    for impl in ("tc", "simt"):
        fv = B200FeatureVolumeManager(h, w, num_depth_bins=D, impl=impl)
        fv.mlp = ref_mlp  # same weights for both kernels (moved by reference like to_fast, cost_volume.py:714)
        fv = fv.cuda()
        v1, fl1, _, m1 = fv(min_depth=mn, max_depth=mx, return_mask=True, **t)
        v3, fl3, _, m3 = fv(min_depth=mn, max_depth=mx, return_mask=True, **one)
        assert torch.equal(v3, v1[1:2]) and torch.equal(fl3, fl1[1:2]) and torch.equal(m3, m1[1:2])
        assert torch.isfinite(v1).all()
        if impl == "tc":
            vtc = v1
    # tensor-core (split-bf16) and strict-fp32 CUDA-core kernels agree to fp32 grade at full size
    assert ((vtc - v1).abs().max() / v1.abs().max()).item() < 1e-4


def test_error_behaviour():
    mgr = B200FeatureVolumeManager(8, 8, num_depth_bins=4, num_source_views=7).cuda()
    t = dev(synthetic.make_volume_inputs(5, 1, 2, 16, 8, 8))
    mn, mx = depth_range()
    with pytest.raises(ValueError):  # the reference fails with a shape error for K != 7
        mgr(min_depth=mn, max_depth=mx, **t)
    mgr2 = B200CostVolumeManager(4, 4, num_depth_bins=4).cuda()
    with pytest.raises(ValueError):
        mgr2(min_depth=mn, max_depth=mx, **t)


@pytest.mark.parametrize("D", [16, 32, 128, 256])
def test_plane_count_sweep_vs_oracle(D):
    """BASELINE config 5: the plane count is a run-time size of both volume kernels (16 ... 256 planes)."""
    B, K, h, w = 1, 7, 24, 32
    inp = synthetic.make_volume_inputs(500 + D, B, K, 16, h, w)
    t = dev(inp)
    mn, mx = depth_range()
    mgr = B200CostVolumeManager(h, w, num_depth_bins=D).cuda()
    cost, lowest, planes_bdhw, _ = mgr(min_depth=mn, max_depth=mx, **t)
    planes = planes_bdhw[0, :, 0, 0].cpu().numpy()
    np.testing.assert_allclose(planes, O.generate_depth_planes(0.25, 5.0, D), rtol=1e-6)
    ref, _, _ = O.cost_volume_dot(inp["cur_feats"], inp["src_feats"], inp["src_extrinsics"], inp["src_Ks"],
                                  inp["cur_invK"], planes)
    assert rel_err(cost.cpu().numpy(), ref) < TOL
    np.testing.assert_array_equal(planes_to_idx(lowest.cpu().numpy(), planes), np.argmax(cost.cpu().numpy(), 1))
    fv = B200FeatureVolumeManager(h, w, num_depth_bins=D, num_source_views=K).cuda()
    torch.manual_seed(D)
    for p in fv.parameters():
        torch.nn.init.normal_(p, std=0.1)
    W = [(fv.mlp.net[i].weight.detach().cpu().numpy(), fv.mlp.net[i].bias.detach().cpu().numpy()) for i in (0, 2, 4)]
    vol, flow, _, mask = fv(min_depth=mn, max_depth=mx, return_mask=True, **t)
    rvol, _, _, rmask = O.feature_volume_mlp(inp["cur_feats"], inp["src_feats"], inp["src_extrinsics"],
                                             inp["src_poses"], inp["src_Ks"], inp["cur_invK"], planes, W)
    assert rel_err(vol.cpu().numpy(), rvol) < TOL
    np.testing.assert_array_equal(planes_to_idx(flow.cpu().numpy(), planes), np.argmax(vol.cpu().numpy(), 1))


def test_per_frame_depth_range_like_the_reference_broadcast():
    """min_depth / max_depth may be [B,1,1,1] tensors (the reference broadcasts them, cost_volume.py:117-126): every
    frame gets its own planes; a wrong element count is an argument error, not a silent first-frame range."""
    B, K, C, h, w, D = 3, 2, 16, 12, 20, 8
    t = dev(synthetic.make_volume_inputs(91, B, K, C, h, w))
    mn = torch.tensor([0.25, 0.5, 1.0], device="cuda").view(B, 1, 1, 1)
    mx = torch.tensor([5.0, 4.0, 8.0], device="cuda").view(B, 1, 1, 1)
    mgr = B200CostVolumeManager(h, w, num_depth_bins=D).cuda()
    cost, lowest, planes_bdhw, _ = mgr(min_depth=mn, max_depth=mx, **t)
    for b in range(B):
        one = {k: v[b:b + 1].contiguous() for k, v in t.items()}
        c1, l1, p1, _ = mgr(min_depth=mn[b:b + 1], max_depth=mx[b:b + 1], **one)
        assert torch.equal(cost[b:b + 1], c1) and torch.equal(lowest[b:b + 1], l1)
        assert torch.equal(planes_bdhw[b, :, 0, 0], p1[0, :, 0, 0])
        np.testing.assert_allclose(p1[0, :, 0, 0].cpu().numpy(),
                                   O.generate_depth_planes(float(mn[b]), float(mx[b]), D), rtol=1e-6)
    with pytest.raises(ValueError):
        mgr(min_depth=mn[:2], max_depth=mx[:2], **t)


@pytest.mark.parametrize("shape", [(4, 7, 96, 128, 64), (2, 8, 37, 53, 10), (1, 1, 9, 7, 1), (1, 3, 120, 160, 96),
                                   (2, 5, 24, 32, 256)])
def test_band_kernel_is_bit_identical_to_gather_kernel(shape):
    """cv_dot_band_kernel (TMA-staged source bands in shared memory) against cv_dot_kernel (per-tap global loads): the
    same arithmetic in the same order, so cost volume and argmax must agree bit for bit -- at the BASELINE cfg2 / cfg4
    shapes, ragged maps, one view / one plane, D not a multiple of 4, and with views that leave the frustum (global and
    skip iterations)."""
    B, K, h, w, D = shape
    inp = synthetic.make_volume_inputs(900 + h, B, K, 16, h, w)
    if K >= 3:  # one view looking backwards (every box off the image), one rotated far enough that boxes do not fit
        flip = np.diag([-1.0, 1.0, -1.0, 1.0]).astype(np.float32)
        inp["src_extrinsics"][:, 0] = flip @ inp["src_extrinsics"][:, 0]
        c, s_ = np.cos(0.6), np.sin(0.6)
        rot = np.array([[c, -s_, 0, 0.3], [s_, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
        inp["src_extrinsics"][:, 1] = rot @ inp["src_extrinsics"][:, 1]
    t = dev(inp)
    mn, mx = depth_range()
    outs = {}
    for impl in ("band", "gather"):
        mgr = B200CostVolumeManager(h, w, num_depth_bins=D, dot_impl=impl).cuda()
        cost, lowest, _, _ = mgr(min_depth=mn, max_depth=mx, **t)
        outs[impl] = (cost, lowest)
    assert torch.equal(outs["band"][0], outs["gather"][0])
    assert torch.equal(outs["band"][1], outs["gather"][1])
