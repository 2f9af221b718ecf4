"""Step time against the number of extra streams of the plans' dependency scheduler (`networks.Plan.N_STREAMS`);
cfg2, CUDA-graph replay, L2 flushed, staged inputs (dev tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from implicit_depth_b200.networks import Plan
from implicit_depth_b200.staging import FrameStaging

torch.set_grad_enabled(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
B, K, H, W, D = 4, 7, 384, 512, 64
st = FrameStaging(B, K, H, W, P=8)
frames = []
for i in range(3):
    cur, src = synthetic.make_frame_batch(7000 + i, B, K, H, W)
    d = st.device_frame("cuda")
    FrameStaging.upload(st.host_frame().fill(cur, src), d)
    frames.append(d)
for n in [int(a) for a in (sys.argv[1:] or ["3", "1", "2", "4", "5", "6", "3"])]:
    Plan.N_STREAMS = n
    m = B200BDModel(default_options(image_width=W, image_height=H, matching_num_depth_bins=D))
    synthetic.init_model_weights(m, seed=0)
    m = m.cuda().eval()
    m.use_cuda_graph = True
    for i in range(4):
        m("test", frames[i % 3].cur, frames[i % 3].src, return_mask=True)
    ts = []
    for i in range(15):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); m("test", frames[i % 3].cur, frames[i % 3].src, return_mask=True); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print(json.dumps({"n_streams": n, "ms_per_forward": round(ts[len(ts) // 2], 3)}), flush=True)
    del m
    torch.cuda.empty_cache()
