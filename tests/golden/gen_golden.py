"""Generate golden vectors by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz

Inputs come from `implicit_depth_b200.synthetic` (numpy PCG64, portable), so the fixtures
only store reference OUTPUTS (+ MLP weights, which depend on torch's RNG).  fp32 outputs of
the reference's slow (loop) managers are the goldens; their fast/efficient twins and an
fp64 run are stored as cross-checks / arbiters.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from implicit_depth_b200 import synthetic  # noqa: E402
from modules import cost_volume as ref_cv  # noqa: E402  (reference)

torch.set_grad_enabled(False)

# name: (seed, B, K, C, h, w, D)
VOLUME_CASES = {
    "cfg1_48x64_k2_d16": (1000, 1, 2, 16, 48, 64, 16),
    "small_24x32_k7_d8": (1001, 2, 7, 16, 24, 32, 8),
    "ragged_20x36_k3_d5": (1002, 1, 3, 16, 20, 36, 5),
    "cfg2_frame_96x128_k7_d64": (2000, 1, 7, 16, 96, 128, 64),
}


def t(x, dtype=torch.float32):
    return torch.from_numpy(np.asarray(x)).to(dtype)


def run_manager(mgr, inp, dtype, return_mask):
    mgr = mgr.to(dtype)
    args = {k: t(v, dtype) for k, v in inp.items()}
    mn = torch.tensor(0.25, dtype=dtype).view(1, 1, 1, 1)
    mx = torch.tensor(5.0, dtype=dtype).view(1, 1, 1, 1)
    cv, lowest, planes, mask = mgr(min_depth=mn, max_depth=mx, return_mask=return_mask, **args)
    return cv, lowest, planes[:, :, 0, 0], mask


def volume_goldens():
    for name, (seed, B, K, C, h, w, D) in VOLUME_CASES.items():
        inp = synthetic.make_volume_inputs(seed, B, K, C, h, w)
        out = {}
        # --- dot-product volume: loop manager (golden) + efficient twin ---
        slow = ref_cv.CostVolumeManager(h, w, num_depth_bins=D)
        cv, lowest, planes, _ = run_manager(slow, inp, torch.float32, False)
        out["dot_cost"] = cv.numpy()
        out["dot_lowest"] = lowest.numpy()
        out["planes"] = planes[0].numpy()
        fast = slow.to_fast()
        cvf, _, _, _ = run_manager(fast, inp, torch.float32, False)
        out["dot_slow_vs_fast_maxabs"] = np.array((cv - cvf).abs().max().item())
        cv64, _, _, _ = run_manager(ref_cv.CostVolumeManager(h, w, num_depth_bins=D), inp, torch.float64, False)
        out["dot_cost_f64"] = cv64.numpy()
        # --- MLP feature volume: loop manager (golden) + fast twin ---
        torch.manual_seed(0)
        fv = ref_cv.FeatureVolumeManager(h, w, num_depth_bins=D, mlp_channels=[202, 128, 128, 1],
                                         matching_dim_size=C, num_source_views=K)
        for i, li in enumerate((0, 2, 4)):
            out[f"mlp_w{i}"] = fv.mlp.net[li].weight.numpy().copy()
            out[f"mlp_b{i}"] = fv.mlp.net[li].bias.numpy().copy()
        vol, lowest, _, mask = run_manager(fv, inp, torch.float32, True)
        out["fv_vol"] = vol.numpy()
        out["fv_lowest"] = lowest.numpy()
        out["fv_mask"] = mask.numpy()
        if h * w * D <= 96 * 128 * 16:
            ff = fv.to_fast()
            volf, _, _, maskf = run_manager(ff, inp, torch.float32, True)
            out["fv_slow_vs_fast_maxabs"] = np.array((vol - volf).abs().max().item())
            out["fv_mask_fast_equal"] = np.array(bool((mask == maskf).all()))
        vol64, _, _, _ = run_manager(fv, inp, torch.float64, True)  # converts fv to fp64 in place
        out["fv_vol_f64"] = vol64.numpy()
        path = os.path.join(HERE, f"volume_{name}.npz")
        if "cfg2" in name:  # keep the fixture small: fp16-free, but drop the fp64 copies to a strided sample
            out["dot_cost_f64"] = out["dot_cost_f64"][:, :, ::4, ::4].copy()
            out["fv_vol_f64"] = out["fv_vol_f64"][:, :, ::4, ::4].copy()
        np.savez_compressed(path, **out)
        print(name, {k: (v.shape, v.dtype) for k, v in out.items() if v.ndim > 0},
              "dot slow-vs-fast", out["dot_slow_vs_fast_maxabs"],
              "fv slow-vs-fast", out.get("fv_slow_vs_fast_maxabs"))


def _sd_sub(sd, prefix):
    return {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


def network_goldens():
    """Reference CVEncoder / BDDecoderPP / SkipDecoder / matching encoder / BinaryMLP with the state dict of
    the seeded B200 containers (identical key names)."""
    import options as ref_options  # reference
    from experiment_modules.bd_model import BDModel  # reference
    from modules import networks as ref_nets  # reference
    from modules import networks_fast as ref_fast  # reference

    from implicit_depth_b200.bd_model import B200BDModel, default_options

    out = {}
    for dec_name in ("unet_pp", "skip"):
        mine = B200BDModel(default_options(image_width=128, image_height=96, matching_num_depth_bins=16,
                                           depth_decoder_name=dec_name))
        out[f"checksum_{dec_name}"] = np.array(synthetic.init_model_weights(mine, seed=0))
        sd = mine.state_dict()
        enc, cv, img = synthetic.make_net_inputs(3000)
        enc_t = [torch.from_numpy(e) for e in enc]
        if dec_name == "unet_pp":
            cve = ref_nets.CVEncoder(16, [48, 64, 160, 256], [64, 128, 256, 384])
            cve.load_state_dict(_sd_sub(sd, "cost_volume_net"))
            cvf = cve(torch.from_numpy(cv), enc_t[1:])
            for i, f in enumerate(cvf):
                out[f"cvenc_{i}"] = f.numpy()
            dec = ref_nets.BDDecoderPP([24, 64, 128, 256, 384])
            dec.load_state_dict(_sd_sub(sd, "depth_decoder"))
            res = dec(enc_t[:1] + cvf)
            for i in range(4):
                out[f"unetpp_s{i}"] = res[f"feature_s{i}_b1hw"].numpy()
            me = ref_nets.ResnetMatchingEncoder(18, 16, pretrained=False)
            me.load_state_dict(_sd_sub(sd, "matching_model"))
            me.eval()
            out["matching"] = me(torch.from_numpy(img)).numpy()
            bm = ref_nets.BinaryMLPNetwork([64, 64, 128, 256], mlp_size=128, use_prior=False)
            bm.load_state_dict(_sd_sub(sd, "binary_mlp"))
            feat = res["feature_s0_b1hw"]
            depth = torch.full((1, 1, 48, 64), 2.5)
            x = torch.cat((depth, feat), 1).permute(0, 2, 3, 1)
            out["binary_pred"] = bm([x], max_scale_only=True)["pred_0"].permute(0, 3, 1, 2).numpy()
        else:
            cve = ref_nets.CVEncoder(16, [48, 64, 160, 256], [64, 128, 256, 384])
            cve.load_state_dict(_sd_sub(sd, "cost_volume_net"))
            cvf = cve(torch.from_numpy(cv), enc_t[1:])
            dec = ref_fast.SkipDecoder([24, 64, 128, 256, 384])
            dec.load_state_dict(_sd_sub(sd, "depth_decoder"))
            res = dec(enc_t[:1] + cvf)
            for i in range(4):
                out[f"skip_s{i}"] = res[f"feature_s{i}_b1hw"].numpy()
    np.savez_compressed(os.path.join(HERE, "nets_96x128.npz"), **out)
    print("nets", {k: v.shape for k, v in out.items()})

    # ---- full BDModel.forward, BASELINE config 1 size (256x192), 7 views (hard-wired), 16 planes ----
    for fv_type, K in (("mlp_feature_volume", 7), ("simple_cost_volume", 2)):
        opts = default_options(image_width=256, image_height=192, matching_num_depth_bins=16,
                               feature_volume_type=fv_type, num_source_views=K)
        mine = B200BDModel(opts)
        checksum = synthetic.init_model_weights(mine, seed=0)
        ro = ref_options.Options()
        ro.image_width, ro.image_height, ro.matching_num_depth_bins = 256, 192, 16
        ro.feature_volume_type = fv_type
        ro.binary_loss_positive_weight = 1.0
        ro.bd_edge_regularision = False
        ref = BDModel(ro)
        sd = dict(mine.state_dict())
        ref.load_state_dict(sd, strict=True)
        ref.eval()
        cur, src = synthetic.make_frame_batch(4000 + K, 1, K, 192, 256)
        cur_t = {k: torch.from_numpy(v) for k, v in cur.items()}
        src_t = {k: torch.from_numpy(v) for k, v in src.items()}
        o = ref("test", cur_t, src_t, unbatched_matching_encoder_forward=True, return_mask=True)
        g = {"checksum": np.array(checksum), "pred_0": o["pred_0"].numpy(), "lowest_cost_bhw": o["lowest_cost_bhw"].numpy()}
        if o["overall_mask_bhw"] is not None:
            g["overall_mask_bhw"] = o["overall_mask_bhw"].numpy()
        np.savez_compressed(os.path.join(HERE, f"model_256x192_{fv_type}.npz"), **g)
        print("model", fv_type, {k: v.shape for k, v in g.items()}, "pred range", g["pred_0"].min(), g["pred_0"].max())


def depth_goldens():
    """Reference DepthModel.forward (depth_model.py:280-440) at 256x192, 7 views, 16 planes, for both
    regression decoders, with the state dict of the seeded B200 container (identical keys)."""
    import options as ref_options  # reference
    from experiment_modules.depth_model import DepthModel  # reference

    from implicit_depth_b200.bd_model import default_options
    from implicit_depth_b200.depth_model import B200DepthModel

    for dec_name in ("unet_pp", "skip"):
        opts = default_options(image_width=256, image_height=192, matching_num_depth_bins=16, depth_decoder_name=dec_name)
        mine = B200DepthModel(opts)
        checksum = synthetic.init_model_weights(mine, seed=0)
        ro = ref_options.Options()
        ro.image_width, ro.image_height, ro.matching_num_depth_bins = 256, 192, 16
        ro.depth_decoder_name = dec_name
        ref = DepthModel(ro)
        ref.load_state_dict(dict(mine.state_dict()), strict=True)
        ref.eval()
        cur, src = synthetic.make_frame_batch(5007, 1, 7, 192, 256)
        cur_t = {k: torch.from_numpy(v) for k, v in cur.items()}
        src_t = {k: torch.from_numpy(v) for k, v in src.items()}
        o = ref("test", cur_t, src_t, unbatched_matching_encoder_forward=True, return_mask=True)
        g = {"checksum": np.array(checksum), "lowest_cost_bhw": o["lowest_cost_bhw"].numpy(),
             "overall_mask_bhw": o["overall_mask_bhw"].numpy()}
        for i in range(4):
            g[f"log_depth_pred_s{i}_b1hw"] = o[f"log_depth_pred_s{i}_b1hw"].numpy()
        np.savez_compressed(os.path.join(HERE, f"depth_model_256x192_{dec_name}.npz"), **g)
        print("depth model", dec_name, {k: v.shape for k, v in g.items()})


def temporal_goldens():
    """use_prior model (implicit_depth_temporal.yaml): prior warp + 66-input binary MLP; and the infer_depth
    bisection of the plain model.  256x192, 7 views, 16 planes, one rendered plane."""
    import options as ref_options  # reference
    from experiment_modules.bd_model import BDModel  # reference
    from utils.geometry_utils import BackprojectDepth  # reference

    from implicit_depth_b200.bd_model import B200BDModel, default_options

    real_cuda = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, *a, **k: self  # BDModel.__init__ calls .cuda() when use_prior (bd_model.py:138)
    try:
        g = {}
        for use_prior in (True, False):
            opts = default_options(image_width=256, image_height=192, matching_num_depth_bins=16, use_prior=use_prior)
            mine = B200BDModel(opts)
            checksum = synthetic.init_model_weights(mine, seed=0)
            ro = ref_options.Options()
            ro.image_width, ro.image_height, ro.matching_num_depth_bins = 256, 192, 16
            ro.use_prior = use_prior
            ro.binary_loss_positive_weight = 1.0
            ro.bd_edge_regularision = False
            ref = BDModel(ro)
            ref.load_state_dict(dict(mine.state_dict()), strict=True)
            ref.eval()
            if use_prior:
                ref.backprojector = BackprojectDepth(96, 128)  # the reference hard-codes 192x256 (bd_model.py:138)
            cur, src = synthetic.make_frame_batch(5000, 1, 7, 192, 256, num_rendered=1, temporal=True)
            cur_t = {k: torch.from_numpy(v) for k, v in cur.items()}
            src_t = {k: torch.from_numpy(v) for k, v in src.items()}
            if use_prior:
                g["checksum_prior"] = np.array(checksum)
                o = ref("test", dict(cur_t), src_t, unbatched_matching_encoder_forward=True, return_mask=True)
                g["prior_pred_0"] = o["pred_0"].numpy()
                g["prior_mask"] = ref.sample_prior(cur_t["rendered_depth"], cur_t["prior_prediction"],
                                                   cur_t["world_T_cam_b44"], cur_t["prior_cam_T_world"],
                                                   cur_t["K_s0_b44"], cur_t["invK_s0_b44"]).numpy()
                no_prior = {k: v for k, v in cur_t.items() if not k.startswith("prior_")}
                o = ref("test", no_prior, src_t, unbatched_matching_encoder_forward=True, return_mask=True)
                g["noprior_pred_0"] = o["pred_0"].numpy()  # first frame of a sequence: prior = -1 (bd_model.py:434)
                # sample_prior alone at the cfg4 size (240x320) with a rendered-depth map that has holes
                ref.backprojector = BackprojectDepth(240, 320)
                c4, _ = synthetic.make_frame_batch(5001, 2, 1, 480, 640, num_rendered=1, temporal=True)
                rng = np.random.default_rng(7)
                rd = rng.uniform(0.5, 4.0, size=c4["rendered_depth"].shape).astype(np.float32)
                rd[rng.uniform(size=rd.shape) < 0.05] = 0.0
                g["cfg4_rendered_depth"] = rd
                t4 = {k: torch.from_numpy(v) for k, v in c4.items()}
                g["cfg4_prior_mask"] = ref.sample_prior(torch.from_numpy(rd), t4["prior_prediction"],
                                                        t4["world_T_cam_b44"], t4["prior_cam_T_world"], t4["K_s0_b44"],
                                                        t4["invK_s0_b44"]).numpy()
            else:
                g["checksum_search"] = np.array(checksum)
                o = ref("test", cur_t, src_t, unbatched_matching_encoder_forward=True, return_mask=True,
                        infer_depth=True)
                g["search_depths"] = o["search_depths"].numpy()
                g["search_pred_0"] = o["pred_0"].numpy()
                # the same search with the evaluation's depth-dependent thresholds (test_bd.py:91-102)
                from utils.binary_metrics_utils import Thresholder  # reference

                real_tcuda = torch.Tensor.cuda
                torch.Tensor.cuda = lambda self, *a, **k: self  # Thresholder.__init__ calls thresholds.cuda()
                try:
                    ref.thresholder = Thresholder(planes=torch.linspace(1.5, 5.0, 8).float(), thresholds=torch.tensor(
                        [0.5, 0.400, 0.3000, 0.3000, 0.3000, 0.3000, 0.300, 0.300]).float())
                finally:
                    torch.Tensor.cuda = real_tcuda
                o = ref("test", cur_t, src_t, unbatched_matching_encoder_forward=True, return_mask=True,
                        infer_depth=True)
                g["thr_bins"] = ref.thresholder.bins.numpy()
                g["thr_vals"] = ref.thresholder.thresholds.numpy()
                g["search_depths_thr"] = o["search_depths"].numpy()
                g["search_pred_0_thr"] = o["pred_0"].numpy()
                ref.thresholder = None
        np.savez_compressed(os.path.join(HERE, "temporal_256x192.npz"), **g)
        print("temporal", {k: v.shape for k, v in g.items()})
    finally:
        torch.nn.Module.cuda = real_cuda


def input_side_goldens():
    """Input side of the step: the reference's own intrinsics pyramid
    (`ScannetDataset.load_intrinsics`, datasets/scannet_dataset.py:435-486, run unbound on files written to a
    scratch directory) and the relative poses its forward hands to the manager (bd_model.py:196-204, captured
    with a forward pre-hook on `BDModel.cost_volume`)."""
    import tempfile
    import types

    # the reference's `datasets/` is a namespace package that the installed HuggingFace `datasets` shadows
    pkg = types.ModuleType("datasets")
    pkg.__path__ = ["/root/reference/datasets"]
    sys.modules["datasets"] = pkg
    from datasets import scannet_dataset as ref_sd  # reference

    import options as ref_options  # reference
    from experiment_modules.bd_model import BDModel  # reference

    from implicit_depth_b200.bd_model import B200BDModel, default_options

    cams = np.array([
        [[577.870605, 0, 319.5, 0], [0, 577.870605, 239.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]],       # ScanNet depth
        [[1170.187988, 0, 647.75, 0], [0, 1170.187988, 483.75, 0], [0, 0, 1, 0], [0, 0, 0, 1]],   # ScanNet colour
        [[886.81, 1.5, 512.0, 0], [0, 890.2, 384.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]],              # skewed pinhole
    ], dtype=np.float64)
    sizes = [(640, 480), (1296, 968), (1024, 768)]
    g = {"K_file": cams, "depth_size": np.array(sizes)}
    with tempfile.TemporaryDirectory() as root:
        Ks, invKs = [], []
        for i, (Kf, (dw, dh)) in enumerate(zip(cams, sizes)):
            scan = f"scene{i:04d}_00"
            os.makedirs(os.path.join(root, scan, "intrinsic"))
            with open(os.path.join(root, scan, f"{scan}.txt"), "w") as f:
                f.write(f"depthWidth = {dw}\ndepthHeight = {dh}\n")
            np.savetxt(os.path.join(root, scan, "intrinsic", "intrinsic_depth.txt"), Kf)
            fake = types.SimpleNamespace(scenes_path=root, include_full_depth_K=False, depth_width=256,
                                         depth_height=192)
            o = ref_sd.ScannetDataset.load_intrinsics(fake, scan, flip=(i == 1))
            Ks.append(np.stack([np.asarray(o[f"K_s{j}_b44"], dtype=np.float32) for j in range(5)]))
            invKs.append(np.stack([np.asarray(o[f"invK_s{j}_b44"], dtype=np.float32) for j in range(5)]))
        g["K_s"] = np.stack(Ks, 1)        # [5 levels, 3 cameras, 4, 4]
        g["invK_s"] = np.stack(invKs, 1)

    opts = default_options(image_width=256, image_height=192, matching_num_depth_bins=16,
                           feature_volume_type="simple_cost_volume", num_source_views=2)
    mine = B200BDModel(opts)
    synthetic.init_model_weights(mine, seed=0)
    ro = ref_options.Options()
    ro.image_width, ro.image_height, ro.matching_num_depth_bins = 256, 192, 16
    ro.feature_volume_type = "simple_cost_volume"
    ro.binary_loss_positive_weight = 1.0
    ro.bd_edge_regularision = False
    ref = BDModel(ro)
    ref.load_state_dict(dict(mine.state_dict()), strict=True)
    ref.eval()
    seen = {}
    ref.cost_volume.register_forward_pre_hook(lambda m, a, kw: seen.update(kw), with_kwargs=True)
    cur, src = synthetic.make_frame_batch(4100, 2, 5, 192, 256)  # B=2, K=5 frames
    ref("test", {k: torch.from_numpy(v) for k, v in cur.items()}, {k: torch.from_numpy(v) for k, v in src.items()},
        unbatched_matching_encoder_forward=True, return_mask=False)
    g["src_cam_T_cur_cam"] = seen["src_extrinsics"].numpy()
    g["cur_cam_T_src_cam"] = seen["src_poses"].numpy()
    np.savez_compressed(os.path.join(HERE, "input_side.npz"), **g)
    print("input side", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    if "--inputs-only" in sys.argv:
        input_side_goldens()
        sys.exit(0)
    if "--temporal-only" in sys.argv:
        temporal_goldens()
        sys.exit(0)
    if "--depth-only" in sys.argv:
        depth_goldens()
        sys.exit(0)
    if "--nets-only" not in sys.argv:
        volume_goldens()
    network_goldens()
    temporal_goldens()
    depth_goldens()
    input_side_goldens()
