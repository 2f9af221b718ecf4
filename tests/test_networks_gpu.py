"""GPU parity of the conv networks and the full BDModel forward: B200 modules (tcgen05 conv path) vs the
reference goldens and the CPU oracle, 1e-3 relative (max|diff| / max|ref|)."""
import numpy as np
import pytest
import torch

from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from oracle import networks as ON
from oracle import planesweep as O

from cases import GOLDEN, argmax_exactness, mask_exactness, record, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def seeded(**kw):
    m = B200BDModel(default_options(**kw))
    checksum = synthetic.init_model_weights(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.cuda().eval(), checksum, sd


def cuda(xs):
    return [torch.from_numpy(x).cuda() for x in xs]


def test_cv_encoder_and_decoders_vs_reference_golden():
    g = np.load(f"{GOLDEN}/nets_96x128.npz")
    enc, cv, img = synthetic.make_net_inputs(3000)
    m, checksum, sd = seeded(image_width=128, image_height=96, matching_num_depth_bins=16)
    golden_ok = abs(checksum - float(g["checksum_unet_pp"])) <= 1e-6 * checksum
    assert golden_ok, "seeded weights differ from the ones the reference goldens were generated with"
    enc_c, cv_c = cuda(enc), torch.from_numpy(cv).cuda()
    cvf = m.cost_volume_net(cv_c, enc_c[1:])
    with torch.no_grad():
        ref_cvf = ON.cv_encoder(sd, "cost_volume_net", torch.from_numpy(cv), [torch.from_numpy(e) for e in enc[1:]])
    for i, f in enumerate(cvf):
        assert rel_err(f.cpu().numpy(), ref_cvf[i].numpy()) < TOL
        if golden_ok:
            assert rel_err(f.cpu().numpy(), g[f"cvenc_{i}"]) < TOL
    dec = m.depth_decoder(enc_c[:1] + cvf)
    with torch.no_grad():
        ref_dec = ON.bd_decoder_pp(sd, "depth_decoder", [torch.from_numpy(enc[0])] + ref_cvf)
    for i in range(4):
        got = dec[f"feature_s{i}_b1hw"].cpu().numpy()
        assert rel_err(got, ref_dec[f"feature_s{i}_b1hw"].numpy()) < TOL
        if golden_ok:
            assert rel_err(got, g[f"unetpp_s{i}"]) < TOL


def test_skip_decoder_vs_reference_golden():
    g = np.load(f"{GOLDEN}/nets_96x128.npz")
    enc, cv, img = synthetic.make_net_inputs(3000)
    m, checksum, sd = seeded(image_width=128, image_height=96, matching_num_depth_bins=16, depth_decoder_name="skip")
    enc_c, cv_c = cuda(enc), torch.from_numpy(cv).cuda()
    cvf = m.cost_volume_net(cv_c, enc_c[1:])
    dec = m.depth_decoder(enc_c[:1] + cvf)
    golden_ok = abs(checksum - float(g["checksum_skip"])) <= 1e-6 * checksum
    assert golden_ok, "seeded weights differ from the ones the reference goldens were generated with"
    with torch.no_grad():
        ref_cvf = ON.cv_encoder(sd, "cost_volume_net", torch.from_numpy(cv), [torch.from_numpy(e) for e in enc[1:]])
        ref = ON.skip_decoder(sd, "depth_decoder", [torch.from_numpy(enc[0])] + ref_cvf)
    for i in range(4):
        got = dec[f"feature_s{i}_b1hw"].cpu().numpy()
        assert rel_err(got, ref[f"feature_s{i}_b1hw"].numpy()) < TOL
        if golden_ok:
            assert rel_err(got, g[f"skip_s{i}"]) < TOL


def test_matching_encoder_vs_reference_and_batch_invariance():
    g = np.load(f"{GOLDEN}/nets_96x128.npz")
    enc, cv, img = synthetic.make_net_inputs(3000)
    m, checksum, sd = seeded(image_width=128, image_height=96, matching_num_depth_bins=16)
    x = torch.from_numpy(img).cuda()
    got = m.matching_model(x)
    with torch.no_grad():
        ref = ON.matching_encoder(sd, "matching_model", torch.from_numpy(img))
    assert rel_err(got.cpu().numpy(), ref.numpy()) < TOL
    assert abs(checksum - float(g["checksum_unet_pp"])) <= 1e-6 * checksum, "seeded weights differ from the goldens'"
    assert rel_err(got.cpu().numpy(), g["matching"]) < TOL
    # the reference runs this encoder one image at a time because batched cuDNN differs in low bits
    # (depth_model.py:235-241); here a frame's features are bit-identical whatever the batch
    rng = np.random.default_rng(1)
    other = torch.from_numpy(rng.standard_normal((2, 3, 96, 128)).astype(np.float32)).cuda()
    both = m.matching_model(torch.cat([other[:1], x, other[1:]], 0))
    assert torch.equal(both[1:2], got)


@pytest.mark.parametrize("fv_type,K", [("mlp_feature_volume", 7), ("simple_cost_volume", 2)])
def test_full_forward_vs_reference_golden(fv_type, K):
    """BASELINE config 1 (256x192, random weights): the whole B200BDModel.forward against the reference
    BDModel run with the same state dict (golden) and the CPU oracle."""
    g = np.load(f"{GOLDEN}/model_256x192_{fv_type}.npz")
    m, checksum, sd = seeded(image_width=256, image_height=192, matching_num_depth_bins=16,
                             feature_volume_type=fv_type, num_source_views=K)
    cur, src = synthetic.make_frame_batch(4000 + K, 1, K, 192, 256)
    cur_c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    src_c = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    out = m("test", cur_c, src_c, unbatched_matching_encoder_forward=True, return_mask=True)
    assert set(out) == {"pred_0", "lowest_cost_bhw", "overall_mask_bhw"}
    assert tuple(out["pred_0"].shape) == (1, 8, 96, 128)
    assert abs(checksum - float(g["checksum"])) <= 1e-6 * checksum, "seeded weights differ from the goldens'"
    assert rel_err(out["pred_0"].cpu().numpy(), g["pred_0"]) < TOL
    _check_indices_and_mask(f"full_forward_256x192/{fv_type}", m, sd, cur, src, out,
                            ref_lowest=g["lowest_cost_bhw"], ref_mask=g["overall_mask_bhw"] if "overall_mask_bhw" in g
                            else None, fv_type=fv_type)
    # CUDA-graph replay gives the same answer as eager launches
    m.use_cuda_graph = True
    out2 = m("test", cur_c, src_c, return_mask=True)
    out3 = m("test", cur_c, src_c, return_mask=True)
    assert torch.equal(out2["pred_0"], out3["pred_0"])
    assert rel_err(out2["pred_0"].cpu().numpy(), out["pred_0"].cpu().numpy()) < 1e-5


def _check_indices_and_mask(case, m, sd, cur, src, out, ref_lowest, ref_mask, fv_type="mlp_feature_volume",
                            oracle_out=None):
    """`lowest_cost_bhw` / `overall_mask_bhw` of a full forward against the reference's (golden or oracle), exact
    outside fp64 near-ties.  The arbiter is the reference algorithm in fp64 on the ORACLE's matching features; the
    product's own features differ from those by fp32 rounding through the encoder, so the tie tolerance is twice the
    measured deviation of the product's volume from the arbiter (recorded; itself bounded by the 1e-3 value bar)."""
    from oracle import planesweep_torch as PT

    opts = m.run_opts
    D, ms = opts.matching_num_depth_bins, opts.matching_scale
    cur_t = {k: torch.from_numpy(v) for k, v in cur.items()}
    src_t = {k: torch.from_numpy(v) for k, v in src.items()}
    if oracle_out is None:
        cpu = B200BDModel(opts)
        cpu.load_state_dict(sd)
        oracle_out = ON.bd_forward(sd, cpu.encoder.eval(), cur_t, src_t, opts, feature_volume=fv_type,
                                   torch_volume=True)
    mf = oracle_out["matching_feats"].double()
    B, K1 = mf.shape[:2]
    h, w = mf.shape[-2:]
    s2c = (src_t["cam_T_world_b44"] @ cur_t["world_T_cam_b44"].unsqueeze(1)).double()
    c2s = (cur_t["cam_T_world_b44"].unsqueeze(1) @ src_t["world_T_cam_b44"]).double()
    Ks, invK = src_t[f"K_s{ms}_b44"].double(), cur_t[f"invK_s{ms}_b44"].double()
    planes64 = PT.depth_planes(opts.min_matching_depth, opts.max_matching_depth, D, torch.float64)
    if fv_type == "mlp_feature_volume":
        W = [(sd[f"cost_volume.mlp.net.{i}.weight"].double(), sd[f"cost_volume.mlp.net.{i}.bias"].double())
             for i in (0, 2, 4)]
        arb = PT.feature_volume_mlp(mf[:, 0], mf[:, 1:], s2c, c2s, Ks, invK, planes64, W, False)[0].numpy()
    else:
        arb = PT.cost_volume_dot(mf[:, 0], mf[:, 1:], s2c, Ks, invK, planes64)[0].numpy()
    planes = O.generate_depth_planes(opts.min_matching_depth, opts.max_matching_depth, D)
    to_idx = lambda z: np.abs(np.log(z)[..., None] - np.log(planes)).argmin(-1)
    # the product's own volume (eager state of the last forward) -> measured deviation from the arbiter
    st = next(iter(m._state.values()))
    vol = st.slots["cv"].cpu().numpy()
    scale = float(arb.max() - arb.min())
    dev = float(np.abs(vol - arb).max()) / scale
    record(case, kind="volume_vs_fp64_arbiter", rel_dev=dev)
    assert dev < TOL
    got_idx = to_idx(out["lowest_cost_bhw"].cpu().numpy())
    np.testing.assert_array_equal(got_idx, np.argmax(vol, 1))  # kernel argmax == first max of its own volume
    n_bad, n_near = argmax_exactness(case, got_idx, to_idx(np.asarray(ref_lowest)), arb, tie_tol=max(1e-5, 2 * dev))
    assert n_bad == n_near, f"{n_bad} plane-index mismatches, only {n_near} at fp64 near-ties"
    if ref_mask is not None:
        edge = PT.mask_edge_distance(s2c, Ks, invK, float(planes64[-1]), h, w).numpy()
        m_bad, m_edge = mask_exactness(case, out["overall_mask_bhw"].cpu().numpy(), np.asarray(ref_mask), edge)
        assert m_bad == m_edge, f"{m_bad} mask flips, only {m_edge} on the window edge"
    return oracle_out


def test_cfg2_benchmark_configuration_vs_oracle():
    """The configuration bench.py times, built the way bench.py builds it -- BASELINE config 2: B=4, 512x384, 7 source
    views, 64 planes, implicit_depth.yaml (mlp_feature_volume + unet_pp), staged inputs read in place, CUDA graph --
    against the CPU oracle frame by frame: pred_0 within 1e-3, plane index and mask exact outside fp64 near-ties
    (reference: cost_volume.py:352-356, bd_model.py:175-311)."""
    from implicit_depth_b200.staging import FrameStaging

    B, K, H, W = 4, 7, 384, 512
    m, _, sd = seeded(image_width=W, image_height=H, matching_num_depth_bins=64)
    m.use_cuda_graph = True
    cur, src = synthetic.make_frame_batch(2000, B, K, H, W)
    staging = FrameStaging(B, K, H, W, P=8, matching_scale=m.run_opts.matching_scale)
    dframe = staging.device_frame("cuda")
    FrameStaging.upload(staging.host_frame().fill(cur, src), dframe)
    out = m("test", dframe.cur, dframe.src, unbatched_matching_encoder_forward=False, return_mask=True)
    out2 = m("test", dframe.cur, dframe.src, unbatched_matching_encoder_forward=False, return_mask=True)  # replay
    assert m._staged_images(dframe.cur, dframe.src, dframe.cur["image_b3hw"]) is not None and len(m._graphs) == 1
    for k in ("pred_0", "lowest_cost_bhw", "overall_mask_bhw"):
        assert torch.equal(out[k], out2[k])
    assert tuple(out["pred_0"].shape) == (B, 8, H // 2, W // 2)
    cpu = B200BDModel(m.run_opts)
    cpu.load_state_dict(sd)
    enc = cpu.encoder.eval()
    worst = 0.0
    for f in range(B):
        cur_f = {k: v[f:f + 1] for k, v in cur.items()}
        src_f = {k: v[f:f + 1] for k, v in src.items()}
        ref = ON.bd_forward(sd, enc, {k: torch.from_numpy(v) for k, v in cur_f.items()},
                            {k: torch.from_numpy(v) for k, v in src_f.items()}, m.run_opts, torch_volume=True)
        e = rel_err(out["pred_0"][f:f + 1].cpu().numpy(), ref["pred_0"].numpy())
        worst = max(worst, e)
        record(f"cfg2_B4_graph_staged/frame{f}", kind="pred_0", rel_err=e)
        assert e < TOL
        # per-frame view of the batch outputs / the batch's volume for the index + mask check
        view = {k: out[k][f:f + 1] for k in ("lowest_cost_bhw", "overall_mask_bhw")}
        st = next(iter(m._state.values()))
        keep = st.slots["cv"]
        st.slots["cv"] = keep[f:f + 1]
        try:
            _check_indices_and_mask(f"cfg2_B4_graph_staged/frame{f}", m, sd, cur_f, src_f, view,
                                    ref_lowest=ref["lowest_cost_bhw"].numpy(),
                                    ref_mask=ref["overall_mask_bhw"].numpy(), oracle_out=ref)
        finally:
            st.slots["cv"] = keep
    record("cfg2_B4_graph_staged", kind="pred_0_worst", rel_err=worst)



@pytest.fixture(autouse=True)
def _strict_fp32_library_convs():
    """The out-of-scope image encoder runs through cuDNN; keep it in strict fp32 so the comparison with the
    CPU golden measures OUR kernels, not cuDNN's TF32 default."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("use_prior", [False, True])
def test_fused_binary_mlp_kernel_vs_oracle(use_prior):
    """csrc/binary_mlp_tc.cu against the numpy restatement of run_mlp_val on a ragged map (pixel count not a
    multiple of the 128-row tile), several planes, with and without the prior column."""
    from implicit_depth_b200.conv import SplitAct
    from implicit_depth_b200.networks import BinaryMLPNetwork, Plan

    rng = np.random.default_rng(11)
    B, H, W, P = 2, 37, 53, 3
    feat = rng.standard_normal((B, 64, H, W)).astype(np.float32)
    depth = rng.uniform(0.5, 6.0, size=(B, P, H, W)).astype(np.float32)
    prior = rng.uniform(-1, 1, size=(B, 1, H, W)).astype(np.float32) if use_prior else None
    torch.manual_seed(3)
    net = BinaryMLPNetwork([64, 64, 128, 256], mlp_size=128, use_prior=use_prior).cuda()
    Wt = [(net.mlps["s0"][i].weight.detach().cpu().numpy(), net.mlps["s0"][i].bias.detach().cpu().numpy())
          for i in (0, 2, 4)]
    g = Plan("cuda")
    fa = SplitAct.from_nchw_torch(torch.from_numpy(feat).cuda())
    d_c = torch.from_numpy(depth).cuda()
    p_c = None if prior is None else torch.from_numpy(prior).cuda()
    pred = net.plan_val(g, fa, lambda: d_c, P, get_prior=(lambda: p_c))
    g.run()
    torch.cuda.synchronize()
    ref = np.concatenate([O.binary_mlp(feat, depth[:, p:p + 1], Wt, prior) for p in range(P)], 1)
    assert rel_err(pred.cpu().numpy(), ref) < 1e-4
    # bisection mode against the numpy loop
    g2 = Plan("cuda")
    search, last = net.plan_search(g2, fa, get_prior=(lambda: p_c))
    g2.run()
    torch.cuda.synchronize()
    z_ref, pred_ref = O.binary_search_depth(feat, Wt, prior)
    assert np.abs(search.cpu().numpy() - z_ref).max() <= 2 * 7.5 / 2**12 + 1e-6
    assert (search.cpu().numpy() != z_ref).mean() < 1e-2
    # depth-dependent thresholds (Thresholder, binary_metrics_utils.py:42-52; test_bd.py:91-102)
    bins = O.thresholder_bins(np.linspace(1.5, 5.0, 8, dtype=np.float32))
    vals = np.array([0.5, 0.4, 0.3, 0.3, 0.3, 0.3, 0.3, 0.3], np.float32)
    thr_c = (torch.from_numpy(bins).cuda(), torch.from_numpy(vals).cuda())
    g3 = Plan("cuda")
    search_t, _ = net.plan_search(g3, fa, get_prior=(lambda: p_c), get_thresholds=(lambda: thr_c))
    g3.run()
    torch.cuda.synchronize()
    z_thr, _ = O.binary_search_depth(feat, Wt, prior, thresholder=(bins, vals))
    assert (z_thr != z_ref).mean() > 0.02  # the thresholds do change the answer on this data
    assert np.abs(search_t.cpu().numpy() - z_thr).max() <= 2 * 7.5 / 2**12 + 1e-6
    assert (search_t.cpu().numpy() != z_thr).mean() < 1e-2


def test_sample_prior_kernel_vs_reference_golden():
    g = np.load(f"{GOLDEN}/temporal_256x192.npz")
    m, _, _ = seeded(image_width=256, image_height=192, matching_num_depth_bins=16, use_prior=True)
    c4, _ = synthetic.make_frame_batch(5001, 2, 1, 480, 640, num_rendered=1, temporal=True)
    t = {k: torch.from_numpy(v).cuda() for k, v in c4.items()}
    rd = torch.from_numpy(g["cfg4_rendered_depth"]).cuda()
    got = m.sample_prior(rd, t["prior_prediction"], t["world_T_cam_b44"], t["prior_cam_T_world"], t["K_s0_b44"],
                         t["invK_s0_b44"]).cpu().numpy()
    assert got.shape == g["cfg4_prior_mask"].shape
    assert (got != g["cfg4_prior_mask"]).mean() < 1e-3  # nearest sampling: rounding-boundary pixels may flip
    ref = O.sample_prior(g["cfg4_rendered_depth"], c4["prior_prediction"], c4["world_T_cam_b44"],
                         c4["prior_cam_T_world"], c4["K_s0_b44"], c4["invK_s0_b44"])
    assert (got != ref).mean() < 1e-3
    assert ((got == -1) >= (g["cfg4_rendered_depth"] <= 0)).all()


@pytest.mark.parametrize("mode", ["prior", "noprior", "search", "search_thr"])
def test_temporal_and_infer_depth_forward_vs_reference_golden(mode):
    """implicit_depth_temporal.yaml path (use_prior: prior warp + 66-input MLP) and the
    infer_depth bisection through B200BDModel.forward, against the reference goldens."""
    g = np.load(f"{GOLDEN}/temporal_256x192.npz")
    use_prior = not mode.startswith("search")
    m, checksum, sd = seeded(image_width=256, image_height=192, matching_num_depth_bins=16, use_prior=use_prior)
    if mode == "search_thr":  # `model.thresholder = Thresholder(...)` of test_bd.py:91-102: anything with these two vectors
        from types import SimpleNamespace

        m.thresholder = SimpleNamespace(bins=torch.from_numpy(g["thr_bins"]).cuda(),
                                        thresholds=torch.from_numpy(g["thr_vals"]).cuda())
    cur, src = synthetic.make_frame_batch(5000, 1, 7, 192, 256, num_rendered=1, temporal=True)
    cur_c = {k: torch.from_numpy(v).cuda() for k, v in cur.items() if mode == "prior" or not k.startswith("prior_")}
    src_c = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    for graph in (False, True):
        m.use_cuda_graph = graph
        out = m("test", dict(cur_c), src_c, return_mask=True, infer_depth=mode.startswith("search"))
        assert abs(checksum - float(g["checksum_prior" if use_prior else "checksum_search"])) <= 1e-6 * checksum, \
            "seeded weights differ from the ones the reference goldens were generated with"
        if mode.startswith("search"):
            want = g["search_depths_thr" if mode == "search_thr" else "search_depths"]
            assert set(out) == {"pred_0", "search_depths", "lowest_cost_bhw", "overall_mask_bhw"}
            sd_got = out["search_depths"].cpu().numpy()
            assert np.abs(sd_got - want).max() <= 4 * 7.5 / 2**12
            assert (np.abs(sd_got - want) > 1e-6).mean() < 2e-2
        else:
            assert rel_err(out["pred_0"].cpu().numpy(), g[f"{mode}_pred_0"]) < TOL


def test_frame_pipeline_matches_direct_forward():
    """The streaming API (H2D / forward / D2H on three streams) returns, batch by batch, what direct calls return."""
    from implicit_depth_b200.pipeline import FramePipeline

    m, _, _ = seeded(image_width=256, image_height=192, matching_num_depth_bins=16)
    m.use_cuda_graph = True
    batches = []
    for i in range(4):
        cur, src = synthetic.make_frame_batch(6000 + i, 1, 7, 192, 256)
        batches.append(({k: torch.from_numpy(v).pin_memory() for k, v in cur.items()},
                        {k: torch.from_numpy(v).pin_memory() for k, v in src.items()}))
    direct = []
    for cur, src in batches:
        o = m("test", {k: v.cuda() for k, v in cur.items()}, {k: v.cuda() for k, v in src.items()}, return_mask=True)
        direct.append({k: v.cpu().clone() for k, v in o.items()})
    pipe = FramePipeline(m, "cuda", return_mask=True)
    got = [{k: v.clone() for k, v in res.items()} for res in pipe.run(iter(batches))]
    assert len(got) == len(direct)
    for g_, d_ in zip(got, direct):
        for k in d_:
            assert torch.equal(g_[k], d_[k]), k
    assert pipe.h2d_bytes > 0 and pipe.d2h_bytes > 0


def test_cfg4_temporal_model_640x480_96_planes_vs_oracle():
    """BASELINE config 4: implicit_depth_temporal.yaml shape (use_prior, 640x480 -> 120x160 matching, 96 planes,
    7 views, one rendered plane, previous prediction warped in): B200BDModel.forward against the CPU oracle run on
    the same inputs and weights, two consecutive frames carrying the prediction forward like test_bd.py:178-214."""
    m, _, sd = seeded(image_width=640, image_height=480, matching_num_depth_bins=96, use_prior=True)
    cpu = B200BDModel(m.run_opts)
    cpu.load_state_dict(sd)
    enc_cpu = cpu.encoder.eval()
    prev_pred = prev_pose = None
    prev_pred_ref = None
    for frame in range(2):
        cur, src = synthetic.make_frame_batch(7000 + frame, 1, 7, 480, 640, num_rendered=1, temporal=True)
        cur_t = {k: torch.from_numpy(v) for k, v in cur.items() if not k.startswith("prior_")}
        src_t = {k: torch.from_numpy(v) for k, v in src.items()}
        cur_c = {k: v.cuda() for k, v in cur_t.items()}
        src_c = {k: v.cuda() for k, v in src_t.items()}
        if prev_pred is not None:
            cur_c["prior_prediction"], cur_c["prior_cam_T_world"] = prev_pred, prev_pose
            cur_t["prior_prediction"], cur_t["prior_cam_T_world"] = prev_pred_ref, prev_pose.cpu()
        out = m("test", cur_c, src_c, return_mask=True)
        ref = ON.bd_forward(sd, enc_cpu, cur_t, src_t, m.run_opts, torch_volume=True)
        assert tuple(out["pred_0"].shape) == (1, 1, 240, 320)
        assert tuple(out["lowest_cost_bhw"].shape) == (1, 120, 160)
        got, want = out["pred_0"].cpu().numpy(), ref["pred_0"].numpy()
        same_prior = np.ones(got.shape, bool)
        if prev_pred is not None:
            # the warped prior is a NEAREST sample (bd_model.py:395-410): a sample position within rounding of a
            # half-integer may pick the neighbouring texel (like an argmax near-tie); such a pixel's MLP input differs,
            # so it is judged by the fp64 arbiter of its sample position and left out of the 1e-3 comparison
            pm_got, pm_ref = cur_c["prior_mask"].cpu().numpy(), ref["prior_mask"].numpy()
            d64 = lambda k: cur_t[k].double().numpy()
            _, xy = O.sample_prior(d64("rendered_depth"), d64("prior_prediction"), d64("world_T_cam_b44"),
                                   d64("prior_cam_T_world"), d64("K_s0_b44"), d64("invK_s0_b44"), return_coords=True)
            to_half = np.minimum(np.abs(xy[:, 0:1] - np.floor(xy[:, 0:1]) - 0.5), np.abs(xy[:, 1:2] - np.floor(xy[:, 1:2]) - 0.5))
            flips = pm_got != pm_ref
            n_flip, n_tie = int(flips.sum()), int((to_half[flips] < 1e-3).sum())
            record(f"cfg4_temporal/frame{frame}", kind="prior_nearest_sample", pixels=int(flips.size), n_bad=n_flip,
                   n_at_rounding_boundary=n_tie, worst_dist_to_half_px=float(to_half[flips].max()) if n_flip else 0.0)
            assert n_flip == n_tie and n_flip < 2e-3 * flips.size
            same_prior = ~flips
        e = float(np.abs(got - want)[same_prior].max() / np.abs(want).max())
        record(f"cfg4_temporal/frame{frame}", kind="pred_0", rel_err=e, rel_err_all_pixels=rel_err(got, want))
        assert e < TOL
        # carry the state forward exactly like the evaluation loop (sigmoid of the logits, test_bd.py:212-214)
        prev_pred = torch.sigmoid(out["pred_0"])
        prev_pred_ref = prev_pred.cpu()  # same state on both sides: the comparison stays per-frame
        prev_pose = cur_c["cam_T_world_b44"]


@pytest.mark.parametrize("layout", ["timm_tf_same", "torchvision"])
def test_native_image_encoder_vs_torch_fp32(layout):
    """EfficientNetV2-S image encoder on the hand-written kernels (image_encoder.py) against the same module evaluated
    by torch on the CPU in fp32 -- the timm `tf_efficientnetv2_s` layout with TF "SAME" padding (the reference's
    encoder, bd_model.py:46-51; even AND odd map sizes: asymmetric (0,1) vs symmetric (1,1) padding at the stride-2
    convs) and the torchvision layout; zero-padded channels must stay exactly zero."""
    from implicit_depth_b200.bd_model import EffNetV2SFeatures
    from implicit_depth_b200.image_encoder import TfEfficientNetV2SFeatures, plan_efficientnet_v2_s
    from implicit_depth_b200.networks import Plan

    enc = (TfEfficientNetV2SFeatures() if layout == "timm_tf_same" else EffNetV2SFeatures()).eval()
    synthetic.init_model_weights(enc, seed=3)
    rng = np.random.default_rng(3100)
    shapes = [(2, 96, 128), (1, 192, 256)] + ([(1, 93, 125)] if layout == "timm_tf_same" else [])
    for (B, H, W) in shapes:
        img = torch.from_numpy(rng.standard_normal((B, 3, H, W)).astype(np.float32))
        with torch.no_grad():
            ref = enc(img)
        g = Plan("cuda")
        slots = {"img": img.cuda()}
        acts = plan_efficientnet_v2_s(g, enc, lambda: slots["img"], B, H, W)
        g.run()
        torch.cuda.synchronize()
        assert [a.Cl for a in acts] == [24, 48, 64, 160, 256]
        for a, r in zip(acts, ref):
            got = a.float_nchw().cpu().numpy()
            assert got.shape[2:] == tuple(r.shape[2:])
            assert rel_err(got[:, :a.Cl], r.numpy()) < TOL
            assert not got[:, a.Cl:].any()


def test_timm_keyed_checkpoint_loads_and_mixed_parity_sizes_are_refused():
    """A state dict with timm's `tf_efficientnetv2_s` key names (what the released checkpoints hold under `encoder.`)
    loads strictly into the default model; a map whose height and width differ in parity at a stride-2 conv needs
    different "SAME" padding per axis, which the kernels do not have: refused loudly, never silently mis-padded."""
    import sys

    from implicit_depth_b200.image_encoder import TfEfficientNetV2SFeatures, plan_efficientnet_v2_s
    from implicit_depth_b200.networks import Plan

    sys.path.insert(0, f"{GOLDEN}/shims")
    import timm  # the shim: an independent restatement of the published model definition (tests only)

    theirs = timm.create_model("tf_efficientnetv2_s_in21ft1k", pretrained=False, features_only=True)
    m = B200BDModel(default_options(image_width=256, image_height=192, matching_num_depth_bins=16))
    missing, unexpected = m.encoder.load_state_dict(theirs.state_dict(), strict=True)
    assert not missing and not unexpected
    assert {k for k in m.state_dict() if k.startswith("encoder.")} == {"encoder." + k for k in theirs.state_dict()}
    g = Plan("cuda")
    with pytest.raises(NotImplementedError):
        plan_efficientnet_v2_s(g, TfEfficientNetV2SFeatures(), lambda: None, 1, 96, 127)


@pytest.mark.parametrize("dec_name", ["unet_pp", "skip"])
def test_depth_model_forward_vs_reference_golden_and_oracle(dec_name):
    """B200DepthModel (regression heads on the same kernels) against the reference DepthModel golden
    and the CPU oracle: log-depth at 4 scales, depth = exp(log-depth), lowest cost, mask."""
    from implicit_depth_b200.depth_model import B200DepthModel

    g = np.load(f"{GOLDEN}/depth_model_256x192_{dec_name}.npz")
    opts = default_options(image_width=256, image_height=192, matching_num_depth_bins=16, depth_decoder_name=dec_name)
    m = B200DepthModel(opts)
    checksum = synthetic.init_model_weights(m, seed=0)
    golden_ok = abs(checksum - float(g["checksum"])) <= 1e-6 * checksum
    assert golden_ok, "seeded weights differ from the ones the reference goldens were generated with"
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    cpu_enc = B200DepthModel(opts).encoder
    cpu_enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    cur, src = synthetic.make_frame_batch(5007, 1, 7, 192, 256)
    ref = ON.depth_forward(sd, cpu_enc.eval(), {k: torch.from_numpy(v) for k, v in cur.items()},
                           {k: torch.from_numpy(v) for k, v in src.items()}, opts)
    m = m.cuda().eval()
    c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    s = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    for graph in (False, True):
        m.use_cuda_graph = graph
        out = m("test", c, s, return_mask=True)
        assert set(out) == {f"{p}depth_pred_s{i}_b1hw" for p in ("log_", "") for i in range(4)} | {
            "lowest_cost_bhw", "overall_mask_bhw"}
        for i in range(4):
            got = out[f"log_depth_pred_s{i}_b1hw"].cpu().numpy()
            assert got.shape == tuple(ref[f"log_depth_pred_s{i}_b1hw"].shape)
            assert rel_err(got, ref[f"log_depth_pred_s{i}_b1hw"].numpy()) < TOL
            if golden_ok:
                assert rel_err(got, g[f"log_depth_pred_s{i}_b1hw"]) < TOL
            np.testing.assert_allclose(out[f"depth_pred_s{i}_b1hw"].cpu().numpy(), np.exp(got), rtol=1e-5)
        assert (out["overall_mask_bhw"].cpu().numpy() != g["overall_mask_bhw"]).mean() < 1e-3


def test_scheduled_plan_is_bit_identical_to_sequential_launches(monkeypatch):
    """The dependency-scheduled launch plans (several streams, per-feature encoder events, SM split, CUDA graph) must
    give exactly the bits of strictly sequential launches on one stream: any missing dependency edge shows up here."""
    cur, src = synthetic.make_frame_batch(4100, 2, 7, 192, 256)
    c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    s = {k: torch.from_numpy(v).cuda() for k, v in src.items()}

    from implicit_depth_b200.networks import Plan

    def run(dag, overlap, graph, n):
        monkeypatch.setattr(Plan, "DAG", dag == "1")
        m, _, _ = seeded(image_width=256, image_height=192, matching_num_depth_bins=16)
        m.overlap_image_encoder = overlap == "1"
        m.use_cuda_graph = graph
        outs = [m("test", c, s, return_mask=True) for _ in range(n)]
        torch.cuda.synchronize()
        return outs

    ref = run("0", "0", False, 1)[0]
    for graph in (False, True):
        for o in run("1", "1", graph, 4):
            for k in ("pred_0", "lowest_cost_bhw", "overall_mask_bhw"):
                assert torch.equal(o[k], ref[k]), f"{k} differs (graph={graph})"


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
def test_sigmoid_upsample_matches_torch(mode):
    """sigmoid_custom + F.interpolate of test_bd.py:225-243 as one kernel, against torch on the GPU."""
    import torch.nn.functional as F

    from implicit_depth_b200 import postprocess

    torch.manual_seed(5)
    for (B, P, h, w, H, W) in ((2, 8, 96, 128, 480, 640), (1, 3, 37, 53, 111, 74), (1, 1, 192, 256, 192, 256)):
        x = torch.randn(B, P, h, w, device="cuda") * 3
        for mult in (1.0, 2.5):
            ref = F.interpolate(1 / (1 + torch.exp(-mult * x)), size=(H, W), mode=mode)
            got = postprocess.sigmoid_upsample(x, (H, W), multiplier=mult, mode=mode)
            assert got.shape == ref.shape and (got - ref).abs().max().item() < 2e-6
        ref = F.interpolate(x, size=(H, W), mode=mode)
        assert (postprocess.resize(x, (H, W), mode=mode) - ref).abs().max().item() < 2e-6


def test_in_place_weight_updates_rebuild_plans_and_graphs():
    """Plans and captured graphs hold packed copies of the weights: any in-place parameter change (sub-module
    load_state_dict, re-initialisation) must be picked up by the next forward, eager or graphed."""
    cur, src = synthetic.make_frame_batch(4200, 1, 7, 192, 256)
    c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    s = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    for graph in (False, True):
        m, _, _ = seeded(image_width=256, image_height=192, matching_num_depth_bins=16)
        m.use_cuda_graph = graph
        a = m("test", c, s, return_mask=True)["pred_0"].clone()
        other = B200BDModel(m.run_opts)
        synthetic.init_model_weights(other, seed=1)
        m.binary_mlp.load_state_dict(other.binary_mlp.state_dict())  # sub-module load: no top-level hook sees it
        b = m("test", c, s, return_mask=True)["pred_0"].clone()
        assert not torch.equal(a, b)
        fresh = B200BDModel(m.run_opts)
        fresh.load_state_dict(m.state_dict())
        want = fresh.cuda().eval()("test", c, s, return_mask=True)["pred_0"]
        assert torch.equal(b, want)


def test_full_forward_is_invariant_to_batch_size_and_composition():
    """A frame's outputs must not depend on how many / which other frames share its batch (the reference protects this
    property by looping its matching encoder, bd_model.py:149-160; here every kernel's reduction order -- including the
    split-K factor of the low-resolution convs -- is a function of the layer geometry only).  Also what makes the
    multi-GPU gather equal to a single-GPU run (tests/test_parallel_gpu.py)."""
    m, _, _ = seeded(image_width=256, image_height=192, matching_num_depth_bins=16)
    cur, src = synthetic.make_frame_batch(4300, 3, 7, 192, 256)
    c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    s = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    full = m("test", c, s, return_mask=True)
    full = {k: v.clone() for k, v in full.items()}
    for sel in ([1], [2, 0]):
        part = m("test", {k: v[sel].contiguous() for k, v in c.items()}, {k: v[sel].contiguous() for k, v in s.items()},
                 return_mask=True)
        for k in ("pred_0", "lowest_cost_bhw", "overall_mask_bhw"):
            assert torch.equal(part[k], full[k][sel]), (k, sel)
