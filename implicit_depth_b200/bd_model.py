"""B200 drop-in for `BDModel.forward` (experiment_modules/bd_model.py:175-311), inference branch.

Same call signature and output dictionary as the reference, same sub-module names and state-dict keys
(`matching_model`, `cost_volume`, `cost_volume_net`, `depth_decoder`, `binary_mlp`, `encoder`), so
`test_bd.py` / `inference/inference.py` can call it unchanged.  Everything on the hot path of SURVEY
section 8 runs as hand-written sm_100a kernels; the EfficientNetV2 image-prior encoder is out of scope
 and stays a PyTorch/cuDNN module that can be injected.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from . import _abi
from .cost_volume import B200CostVolumeManager, B200FeatureVolumeManager
from .networks import BDDecoderPP, BinaryMLPNetwork, CVEncoder, Plan, ResnetMatchingEncoder, SkipDecoder


def default_options(**kw):
    """The fields of the reference's `options.Options` that the forward path reads (options.py:130-160),
    with `configs/models/implicit_depth.yaml` values."""
    o = dict(image_encoder_name="efficientnet", cv_encoder_type="multi_scale_encoder", depth_decoder_name="unet_pp",
             feature_volume_type="mlp_feature_volume", matching_encoder_type="resnet", matching_feature_dims=16,
             matching_scale=1, matching_num_depth_bins=64, min_matching_depth=0.25, max_matching_depth=5.0,
             image_width=512, image_height=384, use_prior=False, num_source_views=7)
    o.update(kw)
    return SimpleNamespace(**o)


class EffNetV2SFeatures(nn.Module):
    """Image-prior encoder stand-in with the channel/stride layout of timm's `tf_efficientnetv2_s_in21ft1k`
    features ([24,48,64,160,256] at /2../32; reference call site bd_model.py:46-51).  Plain torchvision/cuDNN:
    this module is outside the hand-written hot path by design."""

    TAPS = (1, 2, 3, 5, 6)

    def __init__(self):
        super().__init__()
        import torchvision

        net = torchvision.models.efficientnet_v2_s(weights=None)
        self.features = nn.Sequential(*list(net.features)[:7])
        self.num_ch_enc = [24, 48, 64, 160, 256]

    def forward(self, x):
        outs = []
        for i, m in enumerate(self.features):
            x = m(x)
            if i in self.TAPS:
                outs.append(x)
        return outs


class B200BDModel(nn.Module):
    def __init__(self, opts=None, encoder=None):
        super().__init__()
        opts = default_options() if opts is None else opts
        self.run_opts = opts
        if encoder is not None:
            self.encoder = encoder
        elif "efficientnet" in opts.image_encoder_name:
            self.encoder = EffNetV2SFeatures()
        else:
            raise ValueError("Unrecognized option for image encoder type!")
        enc_ch = list(self.encoder.num_ch_enc)
        ms = opts.matching_scale
        if opts.cv_encoder_type != "multi_scale_encoder":
            raise ValueError("Unrecognized option for cost volume encoder type!")
        self.cost_volume_net = CVEncoder(opts.matching_num_depth_bins, enc_ch[ms:], [64, 128, 256, 384])
        dec_in = enc_ch[:ms] + self.cost_volume_net.num_ch_enc
        if opts.depth_decoder_name == "unet_pp":
            self.depth_decoder = BDDecoderPP(dec_in)
        elif opts.depth_decoder_name == "skip":
            self.depth_decoder = SkipDecoder(dec_in)
        else:
            raise ValueError("Unrecognized option for depth decoder name!")
        mh = opts.image_height // (2 ** (ms + 1))
        mw = opts.image_width // (2 ** (ms + 1))
        if opts.feature_volume_type == "simple_cost_volume":
            self.cost_volume = B200CostVolumeManager(mh, mw, num_depth_bins=opts.matching_num_depth_bins)
        elif opts.feature_volume_type == "mlp_feature_volume":
            self.cost_volume = B200FeatureVolumeManager(mh, mw, num_depth_bins=opts.matching_num_depth_bins,
                                                        num_source_views=getattr(opts, "num_source_views", 7))
        else:
            raise ValueError("Unrecognized option for feature volume type!")
        if opts.matching_encoder_type != "resnet":
            raise ValueError("Unrecognized option for matching encoder type!")
        self.matching_model = ResnetMatchingEncoder(18, opts.matching_feature_dims)
        self.binary_mlp = BinaryMLPNetwork(self.depth_decoder.num_ch_dec, mlp_size=128, use_prior=opts.use_prior)
        # training-only buffer of the reference (bd_model.py:100-101), kept so its checkpoints load strictly
        self.bce_loss = nn.Module()
        self.bce_loss.register_buffer("pos_weight", torch.ones(1))
        self._state = {}
        self._graphs = {}
        self.use_cuda_graph = False

    def _apply(self, fn, *a, **k):
        self._state, self._graphs = {}, {}
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------------------------
    def _build(self, B, K, H, W, P, dev):
        """Static launch plans for one input signature."""
        ms = self.run_opts.matching_scale
        D = self.run_opts.matching_num_depth_bins
        slots = {}
        pre = Plan(dev)  # matching encoder
        feats_pm, h, w = self.matching_model.plan(pre, lambda: slots["images"], B * (K + 1), H, W)
        post = Plan(dev)  # cost-volume encoder, decoder, binary MLP
        enc_ch = list(self.encoder.num_ch_enc)
        img_feats = [post.from_f32((lambda i=i: slots["enc"][i]), B, enc_ch[i], H // 2 ** (i + 1), W // 2 ** (i + 1))
                     for i in range(5)]
        cv = post.from_f32(lambda: slots["cv"], B, D, h, w)
        cv_feats = self.cost_volume_net.plan(post, cv, img_feats[ms:])
        dec_in = img_feats[:ms] + cv_feats
        res, _ = self.depth_decoder.plan(post, dec_in, outputs=(0,))
        pred = self.binary_mlp.plan_val(post, res[0], lambda: slots["rendered_depth"], P,
                                        get_prior=(lambda: slots.get("prior")))
        return SimpleNamespace(slots=slots, pre=pre, post=post, feats_pm=feats_pm, h=h, w=w, pred=pred)

    def num_kernel_launches(self, B, K, H, W, P):
        """Hand-written kernel launches per forward for this signature (matching encoder + volume + nets)."""
        st = self._state.get((B, K, H, W, P))
        if st is None:
            return None
        vol = 4 if isinstance(self.cost_volume, B200FeatureVolumeManager) else 2  # prepare, kernel(, argmax)
        return st.pre.n_launches + st.post.n_launches + vol

    @torch.no_grad()
    def _forward_impl(self, cur_image, src_image, src_K, cur_invK, src_cam_T_world, src_world_T_cam, cur_cam_T_world,
                      cur_world_T_cam, rendered_depth, prior, return_mask):
        B, K = src_image.shape[:2]
        H, W = cur_image.shape[-2:]
        P = rendered_depth.shape[1]
        key = (B, K, H, W, P)
        if key not in self._state:
            self._state = {key: self._build(B, K, H, W, P, cur_image.device)}
        st = self._state[key]
        # relative poses, bd_model.py:196-204
        src_cam_T_cur_cam = src_cam_T_world @ cur_world_T_cam.unsqueeze(1)
        cur_cam_T_src_cam = cur_cam_T_world.unsqueeze(1) @ src_world_T_cam
        # image-prior encoder (PyTorch/cuDNN, out of scope)
        st.slots["enc"] = self.encoder(cur_image)
        # matching features for the current + source frames in one batch-invariant pass
        st.slots["images"] = torch.cat([cur_image, src_image.reshape(B * K, 3, H, W)], 0).contiguous()
        st.pre.run()
        N = st.h * st.w
        cur_pm = st.feats_pm[:B]
        src_pm = st.feats_pm[B:].view(B, K, N, -1)
        mn = torch.tensor(self.run_opts.min_matching_depth, device=cur_image.device).view(1, 1, 1, 1) \
            if not hasattr(self, "_mn") or self._mn.device != cur_image.device else self._mn
        mx = torch.tensor(self.run_opts.max_matching_depth, device=cur_image.device).view(1, 1, 1, 1) \
            if not hasattr(self, "_mx") or self._mx.device != cur_image.device else self._mx
        self._mn, self._mx = mn, mx
        cost_volume, lowest_cost, _, overall_mask = self.cost_volume.forward_pixel_major(
            cur_pm, src_pm, src_cam_T_cur_cam, cur_cam_T_src_cam, src_K, cur_invK, mn, mx, None, return_mask, B, K,
            st.h, st.w)
        st.slots["cv"] = cost_volume
        st.slots["rendered_depth"] = rendered_depth
        st.slots["prior"] = prior
        st.post.run()
        return st.pred, lowest_cost, overall_mask

    @torch.no_grad()
    def forward(self, phase, cur_data, src_data, unbatched_matching_encoder_forward=False, return_mask=False,
                infer_depth=False, infer_res=None):
        """Reference signature (bd_model.py:175-184).  `unbatched_matching_encoder_forward` is accepted and
        irrelevant: the matching-encoder kernels are batch-invariant.  Only the inference branch exists."""
        if phase == "train":
            raise NotImplementedError("B200BDModel implements the inference path only")
        if infer_depth:
            raise NotImplementedError("infer_depth (per-pixel binary search) is a next-row item (SURVEY 8f)")
        ms = self.run_opts.matching_scale
        cur_image = cur_data["image_b3hw"]
        _abi.require_cuda(cur_image)
        f = lambda t: t if t.dtype == torch.float32 else t.float()
        args = [f(cur_image).contiguous(), f(src_data["image_b3hw"]), f(src_data[f"K_s{ms}_b44"]),
                f(cur_data[f"invK_s{ms}_b44"]), f(src_data["cam_T_world_b44"]), f(src_data["world_T_cam_b44"]),
                f(cur_data["cam_T_world_b44"]), f(cur_data["world_T_cam_b44"]),
                f(cur_data["rendered_depth"]).contiguous()]
        prior = None
        if self.run_opts.use_prior:
            if cur_data.get("prior_prediction", None) is not None:
                raise NotImplementedError("temporal prior warp (sample_prior) is not built yet")
            prior = -torch.ones_like(args[-1][:, :1]).contiguous()  # bd_model.py:433-434
        if self.use_cuda_graph:
            pred, lowest, mask = self._forward_graphed(args, prior, return_mask)
        else:
            pred, lowest, mask = self._forward_impl(*args, prior, return_mask)
            pred = pred.clone()
        return {"pred_0": pred, "lowest_cost_bhw": lowest, "overall_mask_bhw": mask}

    # ------------------------------------------------------------------------------------
    def _forward_graphed(self, args, prior, return_mask):
        """Whole forward captured once per input signature into a CUDA graph and replayed."""
        key = tuple(tuple(a.shape) for a in args) + (prior is not None, return_mask)
        if key not in self._graphs:
            static = [a.clone() for a in args]
            sprior = None if prior is None else prior.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):  # warm-up: builds plans, packs weights, sets kernel attributes
                    self._forward_impl(*static, sprior, return_mask)
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = self._forward_impl(*static, sprior, return_mask)
            self._graphs = {key: (graph, static, sprior, outs)}
        graph, static, sprior, outs = self._graphs[key]
        for s_, a in zip(static, args):
            s_.copy_(a)
        if prior is not None:
            sprior.copy_(prior)
        graph.replay()
        return tuple(None if o is None else o.clone() for o in outs)
