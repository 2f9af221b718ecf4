"""B200 drop-in for `BDModel.forward` (experiment_modules/bd_model.py:175-311), inference branch.

Same call signature and output dictionary as the reference, same sub-module names and state-dict keys
(`matching_model`, `cost_volume`, `cost_volume_net`, `depth_decoder`, `binary_mlp`, `encoder`), so
`test_bd.py` / `inference/inference.py` can call it unchanged.  Everything on the hot path of SURVEY
section 8 runs as hand-written sm_100a kernels, including the built-in EfficientNetV2-S image-prior encoder
(image_encoder.py); an encoder injected by the caller stays the PyTorch/cuDNN module it is.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn

from . import _abi
from .cost_volume import B200CostVolumeManager, B200FeatureVolumeManager, _Eps, _PixGrid
from .networks import BDDecoderPP, BinaryMLPNetwork, CVEncoder, Plan, ResnetMatchingEncoder, SkipDecoder
from .staging import StagedDict, relative_poses


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
    """torchvision-layout EfficientNetV2-S features ([24,48,64,160,256] at /2../32): same layers as the reference's
    timm encoder but torch-style symmetric padding and `features.N...` keys.  Kept for users who train with
    torchvision weights (`B200BDModel(opts, encoder=EffNetV2SFeatures())` still runs on the hand-written kernels);
    the default encoder is `image_encoder.TfEfficientNetV2SFeatures`, whose keys and padding are the reference's."""

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


def fold_batchnorm(module):
    """Eval-mode copy of `module` with every Conv2d -> BatchNorm2d pair inside an nn.Sequential folded into one
    conv (the out-of-scope cuDNN image encoder spends a fifth of its launches in stand-alone BN kernels)."""
    import copy

    from torch.nn.utils.fusion import fuse_conv_bn_eval

    def walk(m):
        for child in m.children():
            walk(child)
        if isinstance(m, nn.Sequential):
            keys = list(m._modules.keys())
            for a, b in zip(keys, keys[1:]):
                if isinstance(m._modules[a], nn.Conv2d) and isinstance(m._modules[b], nn.BatchNorm2d):
                    m._modules[a] = fuse_conv_bn_eval(m._modules[a], m._modules[b])
                    m._modules[b] = nn.Identity()
        return m

    return walk(copy.deepcopy(module).eval())


def _clone(o):
    if o is None:
        return None
    if isinstance(o, (tuple, list)):
        return tuple(_clone(x) for x in o)
    return o.clone()


class B200BDModel(nn.Module):
    def __init__(self, opts=None, encoder=None):
        super().__init__()
        opts = default_options() if opts is None else opts
        self.run_opts = opts
        # EfficientNetV2-S (the built-in timm-layout module, or either layout passed in) runs on the hand-written
        # kernels (image_encoder.py); any other injected encoder is called as the PyTorch module it is
        from .image_encoder import TfEfficientNetV2SFeatures

        if encoder is not None:
            self.encoder = encoder
        elif "efficientnet" in opts.image_encoder_name:
            # timm `tf_efficientnetv2_s_in21ft1k` layout (bd_model.py:46-51): released checkpoints load as they are
            self.encoder = TfEfficientNetV2SFeatures()
        else:
            raise ValueError("Unrecognized option for image encoder type!")
        self.native_image_encoder = isinstance(self.encoder, (TfEfficientNetV2SFeatures, EffNetV2SFeatures))
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
        if opts.use_prior:
            # buffers of the reference's BackprojectDepth(192, 256) / Project3D (bd_model.py:136-139), kept so that
            # temporal checkpoints load strictly; `sample_prior` computes the same grid in-kernel for any size
            self.backprojector = _PixGrid(192, 256)
            self.projector = _Eps()
        # `model.thresholder = Thresholder(planes, thresholds)` (test_bd.py:91-102): anything with `.bins` and
        # `.thresholds` vectors; None = the fixed 0.5 of bd_model.py:284-285
        self.thresholder = None
        # training-only buffer of the reference (bd_model.py:100-101), kept so its checkpoints load strictly
        self.bce_loss = nn.Module()
        self.bce_loss.register_buffer("pos_weight", torch.ones(1))
        self._state = {}
        self._graphs = {}
        self._versioned = None
        self._thr = None
        self.use_cuda_graph = False
        # the image encoder is independent of the matching encoder + plane sweep until the cost-volume encoder:
        # run it on a side stream so its many small launches overlap the matching encoder and the plane sweep
        self.overlap_image_encoder = True
        # cuDNN's default TF32 convolutions in the image encoder alone push pred_0 to 1.7e-2 of the fp32 reference
        # (scripts/tf32_encoder_check.py; strict fp32: 9e-5), far outside the 1e-3 parity budget: keep it in fp32
        self.encoder_strict_fp32 = True
        self._enc_fast = None
        self._side = None
        # Encoder-ahead mode (pipeline.FramePipeline(encoder_ahead=True)): the image-prior encoder is not part of the
        # forward's graph; `run_encoder` launches it separately (for the NEXT batch, under the back phase of the
        # current one) and the forward starts by copying its outputs into the buffers the cost-volume encoder reads.
        self.encoder_ahead = False
        self._enc_graphs = {}
        self._enc_pending = None       # data_ptr of the images the encoder last ran on and nobody consumed yet
        self.after_encoder_handoff = None  # hook: called once the encoder outputs have been copied out
        # Optional destination views {output key: tensor} (e.g. `parallel.GatherPlan.send_views`): the forward writes
        # its results there instead of into fresh tensors, so a packed gather / download buffer costs no extra copy.
        self.output_views = None

    def _apply(self, fn, *a, **k):
        self._state, self._graphs, self._enc_fast, self._side, self._versioned = {}, {}, None, None, None
        self._thr = None
        self._enc_graphs, self._enc_pending = {}, None
        return super()._apply(fn, *a, **k)

    FRONT_SM_FRACTION = 0.65  # see _front_sm_cap (96 of 148 SMs)
    FV_SM_FRACTION = None     # CTA cap of the plane sweep as a fraction of the SMs; None = the front-end cap

    def _front_sm_cap(self):
        """CTA cap for the persistent kernels of the matching encoder and the plane sweep while the image encoder
        (a long chain of small-grid kernels) runs beside them on its own stream; 0 = no cap."""
        if not (self.native_image_encoder and self.overlap_image_encoder):
            return 0
        if self.encoder_ahead:  # the encoder is not inside this forward
            return 0
        # about two thirds of the machine: measured optimum on B200 (scripts/sm_cap_sweep.py, ms per step at cfg2 with the
        # end-of-round-2 kernels: 86 -> 6.99, 90 -> 6.92, 94 -> 6.82, 98 -> 6.80, 102 -> 6.95, 106 -> 6.95,
        # profiles/r02x_sm_cap_sweep.jsonl; the optimum moved up from 80-90 as the image encoder's chain got shorter)
        return round(self.FRONT_SM_FRACTION *
                     torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count)

    def _run_native_encoder(self, st, dev):
        """Launches the native encoder plan (side stream when `overlap_image_encoder`); returns join()."""
        if not self.overlap_image_encoder:
            st.encp.run()
            return lambda: None
        if self._side is None:
            self._side = _abi.new_stream(dev)
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            st.encp.run()
        return lambda: main.wait_stream(self._side)

    def load_state_dict(self, *a, **k):
        self._state, self._graphs, self._enc_fast = {}, {}, None  # launch plans hold packed copies of the weights
        self._enc_graphs, self._enc_pending = {}, None
        return super().load_state_dict(*a, **k)

    def _run_image_encoder(self, cur_image):
        """Returns (features, join) -- call join() before the features are consumed on the current stream."""
        if self.encoder_strict_fp32:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._run_image_encoder_impl(cur_image)
        return self._run_image_encoder_impl(cur_image)

    def _run_image_encoder_impl(self, cur_image):
        if self.training or not self.overlap_image_encoder:
            return self.encoder(cur_image), (lambda: None)
        key = tuple((p.data_ptr(), p._version) for p in self.encoder.parameters())
        if self._enc_fast is None or self._enc_fast[0] != key:
            self._enc_fast = (key, fold_batchnorm(self.encoder))
        if self._side is None:
            self._side = _abi.new_stream(cur_image.device)
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            feats = self._enc_fast[1](cur_image)
        return feats, (lambda: main.wait_stream(self._side))

    # ------------------------------------------------------------------------------------
    def _build(self, B, K, H, W, P, dev, search=False):
        """Static launch plans for one input signature (P rendered planes, or the infer_depth bisection)."""
        ms = self.run_opts.matching_scale
        D = self.run_opts.matching_num_depth_bins
        slots = {}
        cap = self._front_sm_cap()
        pre = Plan(dev, max_ctas=cap)  # matching encoder (shares the GPU with the image encoder's stream)
        feats_pm, h, w = self.matching_model.plan(pre, lambda: slots["images"], B * (K + 1), H, W,
                                                  feat_layout=self.cost_volume.FEAT_LAYOUT)
        post = Plan(dev)  # cost-volume encoder, decoder, binary MLP
        enc_ch = list(self.encoder.num_ch_enc)
        encp = None
        if self.native_image_encoder:
            from .image_encoder import plan_efficientnet_v2_s

            encp = Plan(dev)  # image-prior encoder on the conv kernels
            img_feats = plan_efficientnet_v2_s(encp, self.encoder, lambda: slots["cur_image"], B, H, W)
            if self.encoder_ahead:
                # the encoder is launched apart from the forward: `post` reads copies of its outputs, so the encoder
                # of the next batch may overwrite its own buffers while this batch is still in the back phase
                enc_out = img_feats
                img_feats = []
                for a in enc_out:
                    c = post.act(a.B, a.H, a.W, a.C)
                    c.Cl = getattr(a, "Cl", a.C)
                    img_feats.append(c)
            else:
                # the encoder runs on its own stream from the start of the forward; every consumer in `post` waits
                # for the one feature map it reads, so the deep encoder stages overlap the first cost-volume-encoder
                # levels
                for a in img_feats:
                    ev = torch.cuda.Event()
                    post.external[id(a)] = ev
                    encp.add((lambda ev=ev: ev.record()), launches=0, reads=[a], writes=[])
        else:
            img_feats = [post.from_f32((lambda i=i: slots["enc"][i]), B, enc_ch[i], H // 2 ** (i + 1),
                                       W // 2 ** (i + 1)) for i in range(5)]
        cv = post.from_f32(lambda: slots["cv"], B, D, h, w)
        cv_feats = self.cost_volume_net.plan(post, cv, img_feats[ms:])
        dec_in = img_feats[:ms] + cv_feats
        pred, search_depths = self._plan_head(post, dec_in, slots, P, search)
        ahead = encp is not None and self.encoder_ahead
        return SimpleNamespace(slots=slots, pre=pre, post=post, encp=encp, feats_pm=feats_pm, h=h, w=w, pred=pred,
                               front_cap=cap,
                               search_depths=search_depths, feat_layout=self.cost_volume.FEAT_LAYOUT,
                               enc_out=(enc_out if ahead else None), post_in=(img_feats if ahead else None),
                               encoder_ahead=ahead)

    def _plan_head(self, post, dec_in, slots, P, search):
        """Decoder + per-pixel binary-occupancy MLP (bd_model.py:260-304).  Returns (pred, search_depths)."""
        res, _ = self.depth_decoder.plan(post, dec_in, outputs=(0,))
        if search:
            search_depths, pred = self.binary_mlp.plan_search(post, res[0], get_prior=(lambda: slots.get("prior")),
                                                              get_thresholds=(lambda: slots.get("thresholds")))
            return pred, search_depths
        pred = self.binary_mlp.plan_val(post, res[0], lambda: slots["rendered_depth"], P,
                                        get_prior=(lambda: slots.get("prior")))
        return pred, None

    def num_kernel_launches(self, B, K, H, W, P, search=False):
        """Hand-written kernel launches per forward for this signature (matching encoder + volume + nets)."""
        st = self._state.get((B, K, H, W, P, search))
        if st is None:
            return None
        vol = 4 if isinstance(self.cost_volume, B200FeatureVolumeManager) else 2  # prepare, kernel(, argmax)
        # + 1: b200_relative_poses
        return st.pre.n_launches + st.post.n_launches + vol + 1 + (st.encp.n_launches if st.encp is not None else 0)

    @torch.no_grad()
    def _forward_impl(self, cur_image, src_image, src_K, cur_invK, src_cam_T_world, src_world_T_cam, cur_cam_T_world,
                      cur_world_T_cam, rendered_depth, prior, return_mask, search=False, images_all=None):
        """`images_all`: the matching encoder's batch [B*(K+1),3,H,W] (current frames first) when the caller staged
        it that way (`staging.FrameStaging`); otherwise it is concatenated here like bd_model.py:162-165."""
        B, K = src_image.shape[:2]
        H, W = cur_image.shape[-2:]
        P = rendered_depth.shape[1]
        st = self._ensure_state(B, K, H, W, P, cur_image.device, search)
        # relative poses, bd_model.py:196-204: both batched products in one launch
        src_cam_T_cur_cam, cur_cam_T_src_cam = relative_poses(src_cam_T_world, src_world_T_cam, cur_cam_T_world,
                                                              cur_world_T_cam)
        # image-prior encoder: native plan on a side stream, or the injected PyTorch module
        if st.encoder_ahead:
            enc_feats, join_encoder = None, (lambda: None)  # launched apart: `run_encoder` + `_encoder_handoff`
        elif st.encp is not None:
            st.slots["cur_image"] = cur_image
            enc_feats, join_encoder = None, self._run_native_encoder(st, cur_image.device)
        else:
            enc_feats, join_encoder = self._run_image_encoder(cur_image)
        # matching features for the current + source frames in one batch-invariant pass
        st.slots["images"] = images_all if images_all is not None else \
            torch.cat([cur_image, src_image.reshape(B * K, 3, H, W)], 0).contiguous()
        st.pre.run()
        N = st.h * st.w
        cur_pm = st.feats_pm[:B]
        src_pm = st.feats_pm[B:].view(B, K, N, -1)
        mn = torch.tensor(self.run_opts.min_matching_depth, device=cur_image.device).view(1, 1, 1, 1) \
            if not hasattr(self, "_mn") or self._mn.device != cur_image.device else self._mn
        mx = torch.tensor(self.run_opts.max_matching_depth, device=cur_image.device).view(1, 1, 1, 1) \
            if not hasattr(self, "_mx") or self._mx.device != cur_image.device else self._mx
        self._mn, self._mx = mn, mx
        self.cost_volume.max_ctas = st.front_cap if (self.FV_SM_FRACTION is None or not st.front_cap) else round(
            self.FV_SM_FRACTION * torch.cuda.get_device_properties(cur_image.device).multi_processor_count)
        try:
            cost_volume, lowest_cost, _, overall_mask = self.cost_volume.forward_pixel_major(
                cur_pm, src_pm, src_cam_T_cur_cam, cur_cam_T_src_cam, src_K, cur_invK, mn, mx, None, return_mask, B, K,
                st.h, st.w)
        finally:
            self.cost_volume.max_ctas = 0
        if st.encp is None:
            join_encoder()
        st.slots["enc"] = enc_feats
        st.slots["cv"] = cost_volume
        st.slots["rendered_depth"] = rendered_depth
        st.slots["prior"] = prior
        st.slots["thresholds"] = self._thr if (search and self.thresholder is not None) else None
        st.post.run()
        if st.encp is not None:
            join_encoder()  # formal join of the side stream (its last op already gates the decoder)
        return st.pred, lowest_cost, overall_mask, st.search_depths

    def _weights_version(self):
        """Changes whenever any parameter / buffer of the model is modified in place (sub-module `load_state_dict`,
        `init_model_weights`, an optimiser step): the launch plans and captured graphs hold packed, BatchNorm-folded
        copies of the weights and must be rebuilt then.  (Tensor versions only grow, so their sum identifies the
        state; re-allocations go through `_apply`.)"""
        if self._versioned is None:
            self._versioned = list(self.parameters()) + list(self.buffers())
        return sum(t._version for t in self._versioned)

    def _sync_weights(self):
        """Drop plans and graphs built from weights that have since been modified (called at the top of forward)."""
        if self._state and next(iter(self._state.values())).weights_version != self._weights_version():
            self._state, self._graphs, self._enc_graphs, self._enc_pending = {}, {}, {}, None

    def _ensure_state(self, B, K, H, W, P, dev, search=False):
        key = (B, K, H, W, P, search)
        st = self._state.get(key)
        wv = self._weights_version()
        if st is not None and st.weights_version != wv:
            st = None
        if st is None or st.feat_layout != self.cost_volume.FEAT_LAYOUT or \
                st.encoder_ahead != (self.encoder_ahead and st.encp is not None):
            self._state = {key: self._build(B, K, H, W, P, dev, search)}
            self._state[key].weights_version = wv
            self._graphs, self._enc_graphs, self._enc_pending = {}, {}, None  # they point into the old plans
        return self._state[key]

    # ---- encoder-ahead mode ------------------------------------------------------------------------------------
    def run_encoder(self, cur_image, K, P, search=False):
        """Launch the image-prior encoder for `cur_image` [B,3,H,W] (fp32, contiguous, stable address: a staging
        slot) on the current stream, as a CUDA graph of its own when `use_cuda_graph`.  Its outputs stay in the
        encoder plan's buffers until the forward of this batch copies them out (`_encoder_handoff`)."""
        if not (self.encoder_ahead and self.native_image_encoder):
            raise RuntimeError("run_encoder needs encoder_ahead mode with the built-in encoder")
        B, _, H, W = cur_image.shape
        self._sync_weights()
        st = self._ensure_state(B, K, H, W, P, cur_image.device, search)
        if not self.use_cuda_graph:
            st.slots["cur_image"] = cur_image
            st.encp.run()
        else:
            key = (B, K, H, W, P, search, cur_image.data_ptr())
            if key not in self._enc_graphs:
                s = _abi.new_stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    st.slots["cur_image"] = cur_image
                    st.encp.run()  # warm-up: packs weights, sets kernel attributes
                torch.cuda.current_stream().wait_stream(s)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    st.encp.run()
                while len(self._enc_graphs) >= self.MAX_STAGED_GRAPHS:
                    self._enc_graphs.pop(next(iter(self._enc_graphs)))
                self._enc_graphs[key] = graph
            self._enc_graphs[key].replay()
        self._enc_pending = cur_image.data_ptr()

    def _encoder_handoff(self, cur_image, K, P, search=False):
        """Start of a forward in encoder-ahead mode: make sure the encoder has run on these images (inline if nobody
        launched it ahead), copy its outputs into the buffers the back phase reads, tell the pipeline."""
        B, _, H, W = cur_image.shape
        st = self._ensure_state(B, K, H, W, P, cur_image.device, search)
        if not st.encoder_ahead:
            return
        if self._enc_pending != cur_image.data_ptr():
            self.run_encoder(cur_image, K, P, search)
        torch._foreach_copy_([a.hi for a in st.post_in] + [a.lo for a in st.post_in],
                             [a.hi for a in st.enc_out] + [a.lo for a in st.enc_out])
        self._enc_pending = None
        if self.after_encoder_handoff is not None:
            self.after_encoder_handoff()

    def sample_prior(self, rendered_depth, prior_prediction, cam_to_world, prior_world_to_cam, K, invK):
        """`BDModel.sample_prior` (bd_model.py:395-410) as one kernel; same argument order as the reference."""
        B, P, H, W = rendered_depth.shape
        if P != 1 or tuple(prior_prediction.shape) != (B, 1, H, W):
            raise ValueError("sample_prior expects one rendered plane and a [B,1,H,W] prior prediction "
                             "(the reference's back-projection broadcasts only for a single plane)")
        _abi.require_cuda(rendered_depth, prior_prediction)
        f = lambda t: (t if t.dtype == torch.float32 else t.float()).contiguous()
        cur_to_prior = torch.matmul(prior_world_to_cam.float(), cam_to_world.float())
        Pm = (K.float() @ cur_to_prior)[:, :3, :].contiguous()
        out = torch.empty((B, 1, H, W), device=rendered_depth.device, dtype=torch.float32)
        _abi.call("b200_sample_prior", _abi.ptr(f(rendered_depth)), _abi.ptr(f(prior_prediction)), _abi.ptr(Pm),
                  _abi.ptr(f(invK)), _abi.ptr(out), B, H, W, _abi.stream_ptr())
        return out

    def _staged_images(self, cur_data, src_data, cur_image):
        """The staged matching-encoder batch when both dictionaries are views of one device `StagedFrame`."""
        frame = getattr(cur_data, "frame", None)
        staged = (isinstance(cur_data, StagedDict) and isinstance(src_data, StagedDict) and frame is not None
                  and src_data.frame is frame and frame.buf.is_cuda
                  and frame.staging.matching_scale == self.run_opts.matching_scale
                  and cur_image.data_ptr() == frame.fields["images"].data_ptr())
        return frame.fields["images"] if staged else None

    @torch.no_grad()
    def forward(self, phase, cur_data, src_data, unbatched_matching_encoder_forward=False, return_mask=False,
                infer_depth=False, infer_res=None):
        """Reference signature (bd_model.py:175-184).  `unbatched_matching_encoder_forward` is accepted and
        irrelevant: the matching-encoder kernels are batch-invariant.  Only the inference branch exists."""
        if phase == "train":
            raise NotImplementedError("B200BDModel implements the inference path only")
        ms = self.run_opts.matching_scale
        cur_image = cur_data["image_b3hw"]
        _abi.require_cuda(cur_image)
        self._sync_weights()
        f = lambda t: t if t.dtype == torch.float32 else t.float()
        args = [f(cur_image).contiguous(), f(src_data["image_b3hw"]), f(src_data[f"K_s{ms}_b44"]),
                f(cur_data[f"invK_s{ms}_b44"]), f(src_data["cam_T_world_b44"]), f(src_data["world_T_cam_b44"]),
                f(cur_data["cam_T_world_b44"]), f(cur_data["world_T_cam_b44"]),
                f(cur_data["rendered_depth"]).contiguous()]
        prior = None
        if self.run_opts.use_prior:
            if cur_data.get("prior_prediction", None) is not None:
                # bd_model.py:423-432 (also stores the warped prior back into the inputs, :432)
                prior = self.sample_prior(args[-1], f(cur_data["prior_prediction"]), f(cur_data["world_T_cam_b44"]),
                                          f(cur_data["prior_cam_T_world"]), f(cur_data["K_s0_b44"]),
                                          f(cur_data["invK_s0_b44"]))
                cur_data["prior_mask"] = prior
            else:
                prior = -torch.ones_like(args[-1][:, :1]).contiguous()  # bd_model.py:433-434
        # a batch staged by `staging.FrameStaging` (one buffer, images already in matching-encoder order): the
        # forward reads the staging slot in place -- no per-tensor copies, no image concatenation
        images_all = self._staged_images(cur_data, src_data, args[0])
        if infer_depth and self.thresholder is not None:
            # static device copies of the Thresholder's vectors (a captured graph keeps reading these addresses)
            bins, vals = f(self.thresholder.bins).reshape(-1), f(self.thresholder.thresholds).reshape(-1)
            if bins.numel() != vals.numel() or bins.numel() == 0:
                raise ValueError("thresholder.bins and thresholder.thresholds must be vectors of equal length")
            if self._thr is None or self._thr[0].numel() != bins.numel() or self._thr[0].device != args[0].device:
                self._thr = (torch.empty(bins.numel(), device=args[0].device),
                             torch.empty(bins.numel(), device=args[0].device))
            self._thr[0].copy_(bins)
            self._thr[1].copy_(vals)
        if self.encoder_ahead and self.native_image_encoder:
            self._encoder_handoff(args[0], args[1].shape[1], args[-1].shape[1], bool(infer_depth))
        if self.use_cuda_graph:
            pred, lowest, mask, search = self._forward_graphed(args, prior, return_mask, bool(infer_depth),
                                                               images_all=images_all)
        else:
            pred, lowest, mask, search = self._forward_impl(*args, prior, return_mask, bool(infer_depth),
                                                            images_all=images_all)
        # results leave the plans' / the graph's static buffers here (one copy each, into `output_views` if given)
        out = {"pred_0": self._emit("pred_0", pred)}
        if infer_depth:
            out["search_depths"] = self._emit("search_depths", search)  # bd_model.py:292
        out["lowest_cost_bhw"] = self._emit("lowest_cost_bhw", lowest)
        out["overall_mask_bhw"] = self._emit("overall_mask_bhw", mask)
        return out

    def _emit(self, name, t):
        if t is None:
            return None
        ov = self.output_views
        if ov is not None and name in ov:
            ov[name].copy_(t)
            return ov[name]
        return t.clone()

    # ------------------------------------------------------------------------------------
    MAX_STAGED_GRAPHS = 8  # one graph per staging slot in use (FramePipeline has two)

    def _forward_graphed(self, args, prior, return_mask, search=False, images_all=None):
        """Whole forward captured once per input signature into a CUDA graph and replayed.  Ordinary inputs are
        copied into the graph's static tensors every call; a staged batch (`images_all` given) is read in place:
        the graph is captured on the staging slot itself, one graph per slot, all sharing the launch plans."""
        thr = self._thr if (search and self.thresholder is not None) else None
        sig = tuple(tuple(a.shape) for a in args) + (prior is not None, return_mask, search,
                                                     None if thr is None else thr[0].data_ptr())
        key = sig + ((images_all.data_ptr(),) if images_all is not None else ())
        if key not in self._graphs:
            static = list(args) if images_all is not None else [a.clone() for a in args]
            sprior = None if prior is None else prior.clone()
            s = _abi.new_stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):  # warm-up: builds plans, packs weights, sets kernel attributes
                    self._forward_impl(*static, sprior, return_mask, search, images_all=images_all)
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = self._forward_impl(*static, sprior, return_mask, search, images_all=images_all)
            # graphs of the same signature (other staging slots, the dictionary path) replay the same launch plans:
            # keep them; any other signature rebuilt the plans, so its graphs are dropped
            keep = {k: v for k, v in self._graphs.items() if k[:len(sig)] == sig}
            while len(keep) >= self.MAX_STAGED_GRAPHS:
                keep.pop(next(iter(keep)))
            keep[key] = (graph, static, sprior, outs, images_all is not None)
            self._graphs = keep
        graph, static, sprior, outs, in_place = self._graphs[key]
        for s_, a in zip(static, args):
            if not in_place or s_.data_ptr() != a.data_ptr():  # (a staged entry the caller replaced is copied in)
                s_.copy_(a)
        if prior is not None:
            sprior.copy_(prior)
        graph.replay()
        return outs
