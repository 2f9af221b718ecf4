"""Steady-state throughput of FramePipeline with and without encoder-ahead overlap (cfg2, staged host frames)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from implicit_depth_b200.pipeline import FramePipeline
from implicit_depth_b200.staging import FrameStaging

torch.set_grad_enabled(False)
B, K, H, W, D = 4, 7, 384, 512, 64
st = FrameStaging(B, K, H, W, P=8)
hosts = []
for i in range(2 if os.environ.get('ONLY_AHEAD') else 3):
    cur, src = synthetic.make_frame_batch(7000 + i, B, K, H, W)
    hosts.append(st.host_frame().fill(cur, src))
N = int(os.environ.get("STEPS", 40))
# settings: "ahead;encoder stream priority;front-end CTA cap (0 = none)" on the command line, e.g.
#   python scripts/time_encoder_ahead.py "0;0;-" "1;0;0" "1;-1;0" "1;-1;120" "1;0;120"
# ("-" = leave the model's default cap rule alone)
SETTINGS = [("1", "0", "-")] if os.environ.get('ONLY_AHEAD') else [("0", "0", "-"), ("1", "0", "-"), ("1", "-1", "-"),
                                                                   ("1", "-1", "120"), ("1", "0", "120"),
                                                                   ("0", "0", "-"), ("1", "0", "-")]
if len(sys.argv) > 1:
    SETTINGS = [tuple(a.split(";")) for a in sys.argv[1:]]
for ahead_s, prio, cap in SETTINGS:
    ahead = ahead_s == "1"
    m = B200BDModel(default_options(image_width=W, image_height=H, matching_num_depth_bins=D))
    if not os.environ.get('NO_INIT'):
        synthetic.init_model_weights(m, seed=0)
    m = m.cuda().eval()
    m.use_cuda_graph = True
    if cap != "-":  # (in encoder-ahead mode the model applies no cap by default)
        frac = int(cap) / torch.cuda.get_device_properties(0).multi_processor_count
        m._front_sm_cap = lambda frac=frac: round(frac * torch.cuda.get_device_properties(0).multi_processor_count)
    pipe = FramePipeline(m, "cuda", encoder_ahead=ahead, encoder_priority=int(prio), return_mask=True)
    feed = lambda n: (hosts[i % len(hosts)] for i in range(n))
    for _ in pipe.run(feed(6)):
        pass
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chk = 0.0
    for res in pipe.run(feed(N)):
        chk += float(res["pred_0"][0, 0, 0, 0])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / N
    print(json.dumps({"encoder_ahead": ahead, "enc_priority": prio, "front_cap": cap, "ms_per_step": round(ms, 3), "frames_per_s": round(1000 * B / ms, 1),
                      "checksum": chk}), flush=True)
    del pipe, m
    torch.cuda.empty_cache()
