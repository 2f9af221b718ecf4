"""CPU ORACLE (test infrastructure, NOT product code) for the convolutional networks and the full
BDModel inference forward.

A functional restatement of the reference's modules on top of the installed torch CPU kernels
(conv2d / interpolate / instance_norm / linear -- the ops the reference itself calls), driven by a plain
state dict with the reference's key names.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import it.

Parity status: PINNED for CVEncoder, BDDecoderPP, SkipDecoder, BinaryMLPNetwork, the matching head and the
full forward: `tests/golden/gen_golden.py` runs the unmodified reference modules from `/root/reference`
with the same state dict and stores their outputs (`tests/golden/nets_*.npz`, `model_*.npz`);
`tests/test_oracle_golden.py` compares.  The antialiased ResNet-18 stem and the EfficientNetV2 encoder
live in third-party packages absent from the reference tree (antialiased-cnns==0.3, timm==0.6.12,
binarydepth_env.yml:26,28): the stem is restated from the published algorithm and is "parity unpinned"
against the original package; the image encoder is an injected module (torchvision stand-in).

Citations are file:line in the reference repository.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import planesweep


def _conv(sd, name, x, stride=1, padding=0, padding_mode="zeros"):
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    if padding_mode == "replicate" and padding > 0:
        x = F.pad(x, [padding] * 4, mode="replicate")
        padding = 0
    return F.conv2d(x, w, b, stride=stride, padding=padding)


def basic_block(sd, p, x, stride=1):
    """`BasicBlock.forward`, modules/layers.py:78-95 (norm = Identity, LeakyReLU(0.2))."""
    out = F.leaky_relu(_conv(sd, p + ".conv1", x, stride, 1), 0.2)
    out = _conv(sd, p + ".conv2", out, 1, 1)
    if (p + ".downsample.0.weight") in sd:
        k = sd[p + ".downsample.0.weight"].shape[-1]
        identity = _conv(sd, p + ".downsample.0", x, stride, k // 2)
    else:
        identity = x
    return F.leaky_relu(out + identity, 0.2)


def cv_encoder(sd, p, x, img_feats):
    """`CVEncoder.forward`, modules/networks.py:208-215."""
    outs = []
    for i in range(len(img_feats)):
        x = basic_block(sd, f"{p}.convs.ds_conv_{i}", x, stride=1 if i == 0 else 2)
        x = torch.cat([x, img_feats[i]], 1)
        x = basic_block(sd, f"{p}.convs.conv_{i}.0", x)
        x = basic_block(sd, f"{p}.convs.conv_{i}.1", x)
        outs.append(x)
    return outs


def upsample(x):
    """`upsample`, utils/generic_utils.py:94-103."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def bd_decoder_pp(sd, p, feats):
    """`BDDecoderPP.forward`, modules/networks.py:64-84 (including its overwrite-per-column behaviour)."""
    prev = list(feats)
    outs = []
    result = {}
    for j in range(1, 5):
        for i in range(4 - j, -1, -1):
            parts = [basic_block(sd, f"{p}.convs.right_conv_{i}{j - 1}", prev[i])]
            parts.append(upsample(basic_block(sd, f"{p}.convs.diag_conv_{i + 1}{j - 1}", prev[i + 1])))
            if i + j != 4:
                parts.append(upsample(basic_block(sd, f"{p}.convs.up_conv_{i + 1}{j}", outs[-1])))
            x = torch.cat(parts, 1)
            x = basic_block(sd, f"{p}.convs.in_conv_{i}{j}.0", x)
            x = basic_block(sd, f"{p}.convs.in_conv_{i}{j}.conv_0", x)
            outs.append(x)
            result[f"feature_s{i}_b1hw"] = x if i == 0 else basic_block(sd, f"{p}.convs.output_{i}.0", x)
        prev = outs[::-1]
    return result


def depth_decoder_pp(sd, p, feats):
    """`DepthDecoderPP.forward`, modules/networks.py:164-183: the UNet++ graph of `bd_decoder_pp` with the heads
    `output_i = (BasicBlock | Identity) -> Conv2d(C, 1, 1)` (:160-163)."""
    feats_out = bd_decoder_pp(sd, p, feats)
    return {f"log_depth_pred_s{i}_b1hw": _conv(sd, f"{p}.convs.output_{i}.1", feats_out[f"feature_s{i}_b1hw"])
            for i in range(4)}


def _conv_block(sd, p, x):
    """`ConvBlock.forward`, modules/networks_fast.py:21-28."""
    x = F.elu(_conv(sd, p + ".conv1", x, 1, 1))
    return F.elu(_conv(sd, p + ".conv2", x, 1, 1))


def skip_decoder(sd, p, feats):
    """`SkipDecoder.forward`, modules/networks_fast.py:83-99."""
    x = feats[-1]
    out = {}
    for n in range(4):
        x = _conv_block(sd, f"{p}.block{n + 1}.pre_concat_conv", x)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = _conv_block(sd, f"{p}.block{n + 1}.post_concat_conv", torch.cat([x, feats[-2 - n]], 1))
        out[f"feature_s{3 - n}_b1hw"] = x
    return out


def skip_decoder_regression(sd, p, feats):
    """`SkipDecoderRegression.forward`, modules/networks_fast.py:137-145 (heads :106-135)."""
    f = skip_decoder(sd, p, feats)
    out = {}
    for n in range(4):
        x = f[f"feature_s{3 - n}_b1hw"]
        x = F.elu(_conv(sd, f"{p}.out{n + 1}.0", x))
        x = F.elu(_conv(sd, f"{p}.out{n + 1}.2", x))
        out[f"log_depth_pred_s{3 - n}_b1hw"] = _conv(sd, f"{p}.out{n + 1}.4", x)
    return out


def depth_forward(sd, encoder, cur, src, opts):
    """`DepthModel.forward(phase="test")`, depth_model.py:280-440 (no flip): the chain of `bd_forward` up to the
    decoder, then the regression decoder and exp()."""
    ms = opts.matching_scale
    cur_image, src_image = cur["image_b3hw"], src["image_b3hw"]
    B, K = src_image.shape[:2]
    src_cam_T_cur_cam = src["cam_T_world_b44"] @ cur["world_T_cam_b44"].unsqueeze(1)
    cur_cam_T_src_cam = cur["cam_T_world_b44"].unsqueeze(1) @ src["world_T_cam_b44"]
    with torch.no_grad():
        enc = encoder(cur_image)
        frames = torch.cat([cur_image.unsqueeze(1), src_image], 1).flatten(0, 1)
        mf = torch.cat([matching_encoder(sd, "matching_model", f[None]) for f in frames], 0)
        mf = mf.view(B, K + 1, *mf.shape[1:])
        from . import planesweep_torch as PT

        tp = PT.depth_planes(opts.min_matching_depth, opts.max_matching_depth, opts.matching_num_depth_bins)
        if opts.feature_volume_type == "mlp_feature_volume":
            W = [(sd[f"cost_volume.mlp.net.{i}.weight"], sd[f"cost_volume.mlp.net.{i}.bias"]) for i in (0, 2, 4)]
            vol, _, lowest, mask = PT.feature_volume_mlp(mf[:, 0], mf[:, 1:], src_cam_T_cur_cam, cur_cam_T_src_cam,
                                                         src[f"K_s{ms}_b44"], cur[f"invK_s{ms}_b44"], tp, W, True)
        else:
            vol, _, lowest = PT.cost_volume_dot(mf[:, 0], mf[:, 1:], src_cam_T_cur_cam, src[f"K_s{ms}_b44"],
                                                cur[f"invK_s{ms}_b44"], tp)
            mask = None
        cvf = cv_encoder(sd, "cost_volume_net", vol, enc[ms:])
        feats = list(enc[:ms]) + cvf
        if opts.depth_decoder_name == "unet_pp":
            out = depth_decoder_pp(sd, "depth_decoder", feats)
        else:
            out = skip_decoder_regression(sd, "depth_decoder", feats)
        for k in list(out.keys()):
            out[k.replace("log_", "")] = torch.exp(out[k])  # depth_model.py:426-435
    out["lowest_cost_bhw"] = lowest
    out["overall_mask_bhw"] = mask
    return out


def _bn(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, eps)


def matching_encoder(sd, p, img):
    """`ResnetMatchingEncoder.forward`, modules/networks.py:264-287, eval mode.  Stem = antialiased-cnns 0.3
    resnet18: conv7x7/2, BN, ReLU, MaxPool(2, s1) + BlurPool(filt 4, s2, reflect pad (1,2,1,2)), layer1."""
    x = F.relu(_bn(sd, p + ".net.1", F.conv2d(img, sd[p + ".net.0.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 2, 1)
    x = F.conv2d(F.pad(x, [1, 2, 1, 2], mode="reflect"), sd[p + ".net.3.1.filt"], stride=2, groups=x.shape[1])
    for b in range(2):
        q = f"{p}.net.4.{b}"
        h = F.relu(_bn(sd, q + ".bn1", F.conv2d(x, sd[q + ".conv1.weight"], None, 1, 1)))
        h = _bn(sd, q + ".bn2", F.conv2d(h, sd[q + ".conv2.weight"], None, 1, 1))
        x = F.relu(h + x)
    x = _conv(sd, p + ".net.5", x)
    x = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)
    x = _conv(sd, p + ".net.8", x, 1, 1, "replicate")
    return F.instance_norm(x, eps=1e-5)


def binary_mlp_val(sd, p, feat, rendered_depth, prior=None):
    """`BDModel.run_mlp_val` looped over rendered planes, bd_model.py:293-304, 412-442."""
    W = [(sd[f"{p}.mlps.s0.{i}.weight"].numpy(), sd[f"{p}.mlps.s0.{i}.bias"].numpy()) for i in (0, 2, 4)]
    outs = []
    for idx in range(rendered_depth.shape[1]):
        pr = None if prior is None else prior.numpy()
        outs.append(planesweep.binary_mlp(feat.numpy(), rendered_depth[:, idx:idx + 1].numpy(), W, pr))
    return torch.from_numpy(np.concatenate(outs, 1))


def bd_forward(sd, encoder, cur, src, opts, feature_volume="mlp_feature_volume", decoder="unet_pp", return_mask=True,
               torch_volume=False, infer_depth=False, thresholder=None):
    """`BDModel.forward(phase="test")`, bd_model.py:175-311.  cur/src: dicts of CPU float tensors; `encoder`:
    the image-prior module (same instance the product uses, on CPU)."""
    ms = opts.matching_scale
    cur_image, src_image = cur["image_b3hw"], src["image_b3hw"]
    B, K = src_image.shape[:2]
    src_cam_T_cur_cam = src["cam_T_world_b44"] @ cur["world_T_cam_b44"].unsqueeze(1)  # :200
    cur_cam_T_src_cam = cur["cam_T_world_b44"].unsqueeze(1) @ src["world_T_cam_b44"]  # :204
    with torch.no_grad():
        enc = encoder(cur_image)  # :218
        frames = torch.cat([cur_image.unsqueeze(1), src_image], 1).flatten(0, 1)
        mf = torch.cat([matching_encoder(sd, "matching_model", f[None]) for f in frames], 0)  # :149-160
        mf = mf.view(B, K + 1, *mf.shape[1:])
        cur_f, src_f = mf[:, 0].numpy(), mf[:, 1:].numpy()
        planes = planesweep.generate_depth_planes(opts.min_matching_depth, opts.max_matching_depth,
                                                  opts.matching_num_depth_bins)
        a = (cur_f, src_f, src_cam_T_cur_cam.numpy(), cur_cam_T_src_cam.numpy(), src[f"K_s{ms}_b44"].numpy(),
             cur[f"invK_s{ms}_b44"].numpy())
        if torch_volume:  # multi-threaded torch-CPU port (the timed CPU baseline)
            from . import planesweep_torch as PT

            tp = PT.depth_planes(opts.min_matching_depth, opts.max_matching_depth, opts.matching_num_depth_bins)
            ta = [torch.from_numpy(x) for x in a]
            if feature_volume == "mlp_feature_volume":
                W = [(sd[f"cost_volume.mlp.net.{i}.weight"], sd[f"cost_volume.mlp.net.{i}.bias"]) for i in (0, 2, 4)]
                vol, _, lowest, mask = PT.feature_volume_mlp(ta[0], ta[1], ta[2], ta[3], ta[4], ta[5], tp, W,
                                                             return_mask)
                mask = None if mask is None else mask.numpy()
            else:
                vol, _, lowest = PT.cost_volume_dot(ta[0], ta[1], ta[2], ta[4], ta[5], tp)
                mask = None
            vol, lowest = vol.numpy(), lowest.numpy()
        elif feature_volume == "mlp_feature_volume":
            W = [(sd[f"cost_volume.mlp.net.{i}.weight"].numpy(), sd[f"cost_volume.mlp.net.{i}.bias"].numpy())
                 for i in (0, 2, 4)]
            vol, _, lowest, mask = planesweep.feature_volume_mlp(a[0], a[1], a[2], a[3], a[4], a[5], planes, W,
                                                                 return_mask)
        else:
            vol, _, lowest = planesweep.cost_volume_dot(a[0], a[1], a[2], a[4], a[5], planes)
            mask = None
        cvf = cv_encoder(sd, "cost_volume_net", torch.from_numpy(vol), enc[ms:])  # :254-257
        feats = list(enc[:ms]) + cvf  # :258
        dec = bd_decoder_pp(sd, "depth_decoder", feats) if decoder == "unet_pp" else skip_decoder(sd, "depth_decoder",
                                                                                                 feats)
        prior = None
        if getattr(opts, "use_prior", False):
            if cur.get("prior_prediction", None) is not None:  # :423-431
                prior = torch.from_numpy(planesweep.sample_prior(
                    cur["rendered_depth"].numpy(), cur["prior_prediction"].numpy(), cur["world_T_cam_b44"].numpy(),
                    cur["prior_cam_T_world"].numpy(), cur["K_s0_b44"].numpy(), cur["invK_s0_b44"].numpy()))
            else:
                prior = -torch.ones_like(cur["rendered_depth"][:, :1])  # :433-434
        search = None
        if infer_depth:  # :273-292
            W = [(sd[f"binary_mlp.mlps.s0.{i}.weight"].numpy(), sd[f"binary_mlp.mlps.s0.{i}.bias"].numpy())
                 for i in (0, 2, 4)]
            z, pred = planesweep.binary_search_depth(dec["feature_s0_b1hw"].numpy(), W,
                                                     None if prior is None else prior.numpy(), thresholder=thresholder)
            search, pred = torch.from_numpy(z), torch.from_numpy(pred)
        else:
            pred = binary_mlp_val(sd, "binary_mlp", dec["feature_s0_b1hw"], cur["rendered_depth"], prior)
    return {"pred_0": pred, "search_depths": search, "prior_mask": prior, "lowest_cost_bhw": torch.from_numpy(lowest),
            "overall_mask_bhw": None if mask is None else torch.from_numpy(mask), "cost_volume": torch.from_numpy(vol),
            "feature_s0": dec["feature_s0_b1hw"], "matching_feats": mf}
