"""B200 drop-in for `DepthModel.forward` (experiment_modules/depth_model.py:280-440), inference branch -- the
regression sibling of the binary-depth model (`test_reg.py:70-72,162-168`).

Identical chain up to the decoder (image-prior encoder, matching encoder, plane-sweep volume, cost-volume encoder: the
same launch plans as `B200BDModel`); the decoder is `DepthDecoderPP` (modules/networks.py:118-183) or
`SkipDecoderRegression` (modules/networks_fast.py:102-145), whose `output_*` heads are live here, followed by the
exp() of depth_model.py:426-435.  Same sub-module names and state-dict keys as the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _abi
from .bd_model import B200BDModel, default_options
from .cost_volume import _Eps, _PixGrid
from .networks import DepthDecoderPP, SkipDecoderRegression


class B200DepthModel(B200BDModel):
    def __init__(self, opts=None, encoder=None):
        opts = default_options() if opts is None else opts
        super().__init__(opts, encoder=encoder)
        ms = opts.matching_scale
        dec_in = list(self.encoder.num_ch_enc)[:ms] + self.cost_volume_net.num_ch_enc
        if opts.depth_decoder_name == "unet_pp":
            self.depth_decoder = DepthDecoderPP(dec_in)  # depth_model.py:163-164
        elif opts.depth_decoder_name == "skip":
            self.depth_decoder = SkipDecoderRegression(dec_in)  # :165-166
        else:
            raise ValueError("Unrecognized option for depth decoder name!")
        # modules of the binary model that the regression model does not have
        del self.binary_mlp, self.bce_loss
        for name in ("backprojector", "projector"):
            if hasattr(self, name):
                delattr(self, name)
        # buffers of the reference's loss helpers (depth_model.py:173-190), kept so its checkpoints load strictly
        h2, w2 = opts.image_height // 2, opts.image_width // 2
        self.mv_depth_loss = nn.Module()
        self.mv_depth_loss.backproject = _PixGrid(h2, w2)
        self.mv_depth_loss.project = _Eps()
        self.compute_normals = nn.Module()
        self.compute_normals.backproject = _PixGrid(h2, w2)

    def _plan_head(self, post, dec_in, slots, P, search):
        outs = self.depth_decoder.plan_depth(post, dec_in)
        flat = tuple(t for i in range(4) for t in outs[i])  # (log_s0, depth_s0, log_s1, depth_s1, ...)
        return flat, None

    @torch.no_grad()
    def forward(self, phase, cur_data, src_data, unbatched_matching_encoder_forward=False, return_mask=False):
        """Reference signature (depth_model.py:280-287).  Returns `log_depth_pred_s{i}_b1hw`, `depth_pred_s{i}_b1hw`
        (i = 0..3), `lowest_cost_bhw`, `overall_mask_bhw`.  Only the inference branch exists (no flip augmentation)."""
        if phase == "train":
            raise NotImplementedError("B200DepthModel implements the inference path only")
        ms = self.run_opts.matching_scale
        cur_image = cur_data["image_b3hw"]
        _abi.require_cuda(cur_image)
        self._sync_weights()
        f = lambda t: t if t.dtype == torch.float32 else t.float()
        B = cur_image.shape[0]
        no_planes = torch.empty((B, 0, 1, 1), device=cur_image.device, dtype=torch.float32)
        args = [f(cur_image).contiguous(), f(src_data["image_b3hw"]), f(src_data[f"K_s{ms}_b44"]),
                f(cur_data[f"invK_s{ms}_b44"]), f(src_data["cam_T_world_b44"]), f(src_data["world_T_cam_b44"]),
                f(cur_data["cam_T_world_b44"]), f(cur_data["world_T_cam_b44"]), no_planes]
        images_all = self._staged_images(cur_data, src_data, args[0])
        if self.encoder_ahead and self.native_image_encoder:
            self._encoder_handoff(args[0], args[1].shape[1], 0, False)
        if self.use_cuda_graph:
            pred, lowest, mask, _ = self._forward_graphed(args, None, return_mask, False, images_all=images_all)
        else:
            pred, lowest, mask, _ = self._forward_impl(*args, None, return_mask, False, images_all=images_all)
        out = {}
        for i in range(4):
            out[f"log_depth_pred_s{i}_b1hw"] = self._emit(f"log_depth_pred_s{i}_b1hw", pred[2 * i])
            out[f"depth_pred_s{i}_b1hw"] = self._emit(f"depth_pred_s{i}_b1hw", pred[2 * i + 1])
        out["lowest_cost_bhw"] = self._emit("lowest_cost_bhw", lowest)
        out["overall_mask_bhw"] = self._emit("overall_mask_bhw", mask)
        return out
