"""The reference's algorithm in plain PyTorch / cuDNN on the same GPU (evidence for the north-star target ">= 4x the
reference's own PyTorch/cuDNN forward on 1 x B200"; dev tool, its output is committed under profiles/).

The reference itself (/root/reference) cannot travel to the GPU box; this runs the op-for-op torch restatement of
`oracle/` (the per-plane loop of FeatureVolumeManager.build_cost_volume, cost_volume.py:437-706, i.e. test_bd.py
without --fast_cost_volume; conv nets through F.conv2d = cuDNN) on CUDA tensors, B=4 at cfg2, with cuDNN/cuBLAS TF32
on (PyTorch's default, what the reference would run) and off (the strict-fp32 numerics the parity bar is defined on).
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from oracle import networks as ON, planesweep_torch as PT

torch.set_grad_enabled(False)
B, K, H, W, D = 4, 7, 384, 512, 64
opts = default_options(image_width=W, image_height=H, matching_num_depth_bins=D)
model = B200BDModel(opts)
synthetic.init_model_weights(model, seed=0)
sd = {k: v.detach().cuda() for k, v in model.state_dict().items()}
enc = model.encoder.cuda().eval()
cur, src = synthetic.make_frame_batch(2000, B, K, H, W)
cur = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
src = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
planes = PT.depth_planes(opts.min_matching_depth, opts.max_matching_depth, D).cuda()
Wm = [(sd[f"cost_volume.mlp.net.{i}.weight"], sd[f"cost_volume.mlp.net.{i}.bias"]) for i in (0, 2, 4)]


def forward(unbatched):
    ms = opts.matching_scale
    extr = src["cam_T_world_b44"] @ cur["world_T_cam_b44"].unsqueeze(1)
    poses = cur["cam_T_world_b44"].unsqueeze(1) @ src["world_T_cam_b44"]
    feats = enc(cur["image_b3hw"])
    frames = torch.cat([cur["image_b3hw"].unsqueeze(1), src["image_b3hw"]], 1).flatten(0, 1)
    if unbatched:  # test_bd.py default (bd_model.py:149-160)
        mf = torch.cat([ON.matching_encoder(sd, "matching_model", f[None]) for f in frames], 0)
    else:
        mf = ON.matching_encoder(sd, "matching_model", frames)
    mf = mf.view(B, K + 1, *mf.shape[1:])
    vol, _, lowest, mask = PT.feature_volume_mlp(mf[:, 0], mf[:, 1:].contiguous(), extr, poses, src[f"K_s{ms}_b44"],
                                                 cur[f"invK_s{ms}_b44"], planes, Wm, True)
    cvf = ON.cv_encoder(sd, "cost_volume_net", vol, feats[ms:])
    dec = ON.bd_decoder_pp(sd, "depth_decoder", list(feats[:ms]) + cvf)
    feat, rd = dec["feature_s0_b1hw"], cur["rendered_depth"]
    outs = []
    for idx in range(rd.shape[1]):  # BDModel.run_mlp_val per rendered plane (bd_model.py:293-304, 412-442)
        x = torch.cat([rd[:, idx:idx + 1], feat], 1).permute(0, 2, 3, 1)
        for i in (0, 2, 4):
            x = F.linear(x, sd[f"binary_mlp.mlps.s0.{i}.weight"], sd[f"binary_mlp.mlps.s0.{i}.bias"])
            if i < 4:
                x = F.elu(x)
        outs.append(x.permute(0, 3, 1, 2))
    return torch.cat(outs, 1)


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]


out = {"config": "cfg2: B=4, 512x384, K=7, D=64, mlp_feature_volume (per-plane loop), unet_pp; torch %s" % torch.__version__}
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    for unb in (True, False):
        ms_ = timeit(lambda: forward(unb))
        out[f"tf32={tf32},unbatched_matching={unb}"] = {"ms_per_batch": ms_, "frames_per_s": 1000.0 * B / ms_}
print(json.dumps(out, indent=1))
