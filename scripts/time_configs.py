"""Forward latency of the other BASELINE configurations / model variants (dev tool; CUDA-graph replay, L2 flushed)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from implicit_depth_b200.depth_model import B200DepthModel

torch.set_grad_enabled(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=10):
    for _ in range(4): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]


def run(name, cls, B, H, W, D, infer_depth=False, temporal=False, **kw):
    opts = default_options(image_width=W, image_height=H, matching_num_depth_bins=D, use_prior=temporal, **kw)
    m = cls(opts)
    synthetic.init_model_weights(m, seed=0)
    m = m.cuda().eval(); m.use_cuda_graph = True
    cur, src = synthetic.make_frame_batch(7000, B, 7, H, W)
    c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    s = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    if temporal:
        c["rendered_depth"] = c["rendered_depth"][:, :1].contiguous()
        c["prior_prediction"] = torch.rand(B, 1, H // 2, W // 2, device="cuda")
        c["prior_cam_T_world"] = c["cam_T_world_b44"].clone()
    if cls is B200DepthModel:
        fn = lambda: m("test", c, s, return_mask=True)
    else:
        fn = lambda: m("test", c, s, return_mask=True, infer_depth=infer_depth)
    ms = timeit(fn)
    print(json.dumps({"config": name, "B": B, "ms_per_forward": ms, "frames_per_s": 1000.0 * B / ms}), flush=True)


run("cfg2 BD mlp_feature_volume unet_pp (bench)", B200BDModel, 4, 384, 512, 64)
run("cfg2 BD simple_cost_volume unet_pp", B200BDModel, 4, 384, 512, 64, feature_volume_type="simple_cost_volume")
run("cfg2 BD mlp_feature_volume skip decoder", B200BDModel, 4, 384, 512, 64, depth_decoder_name="skip")
run("cfg2 BD infer_depth bisection", B200BDModel, 4, 384, 512, 64, infer_depth=True)
run("cfg2 DepthModel unet_pp (8f row 1)", B200DepthModel, 4, 384, 512, 64)
run("cfg4 temporal BD 640x480 D=96 B=1", B200BDModel, 1, 480, 640, 96, temporal=True)
run("cfg1 BD 256x192 D=16 B=1", B200BDModel, 1, 192, 256, 16)
