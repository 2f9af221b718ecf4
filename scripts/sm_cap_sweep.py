"""SM split between the image-prior encoder (side stream) and the matching encoder / plane sweep (dev tool).
Sweeps the CTA cap of the front-end stages (`B200BDModel.FRONT_SM_FRACTION`); cfg2, CUDA-graph replay, L2 flushed,
staged inputs.  One JSON line per setting."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from implicit_depth_b200.staging import FrameStaging

torch.set_grad_enabled(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
B, K, H, W, D = 4, 7, 384, 512, 64
m = B200BDModel(default_options(image_width=W, image_height=H, matching_num_depth_bins=D))
synthetic.init_model_weights(m, seed=0)
m = m.cuda().eval()
m.use_cuda_graph = True
st = FrameStaging(B, K, H, W, P=8)
frames = []
for i in range(3):
    cur, src = synthetic.make_frame_batch(7000 + i, B, K, H, W)
    d = st.device_frame("cuda")
    FrameStaging.upload(st.host_frame().fill(cur, src), d)
    frames.append(d)
torch.cuda.synchronize()


def timeit(n=15):
    for i in range(4):
        m("test", frames[i % 3].cur, frames[i % 3].src, return_mask=True)
    ts = []
    for i in range(n):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); m("test", frames[i % 3].cur, frames[i % 3].src, return_mask=True); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]


# front-end CTA cap (matching encoder [+ plane sweep] while the image encoder runs beside them; 0 = no cap) and, after
# a colon, the plane sweep's own cap: "90" or "90:120"
SETTINGS = sys.argv[1:] or ["0", "80", "90", "100", "90:110", "90:130", "90:148", "100:148", "110:148"]
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
for spec in SETTINGS:
    front, _, fv = spec.partition(":")
    m.FRONT_SM_FRACTION = int(front) / n_sm
    m.FV_SM_FRACTION = (int(fv) / n_sm) if fv else None
    m._state, m._graphs = {}, {}  # the plans bake the cap
    torch.cuda.synchronize()
    ms = timeit()
    print(json.dumps({"front_cap": int(front), "fv_cap": int(fv) if fv else int(front), "ms_per_forward": round(ms, 3),
                      "frames_per_s": round(1000.0 * B / ms, 1)}), flush=True)
