"""SM split between the image-prior encoder (side stream) and the matching encoder / plane sweep (dev tool).
Sweeps the CTA caps of the two front-end stages and the priority of the encoder stream; cfg2, CUDA-graph replay,
L2 flushed, staged inputs.  One JSON line per setting."""
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


# (matching-encoder cap, plane-sweep cap, encoder stream priority, plane-sweep split "frames:cap,...")
SETTINGS = [(74, 74, 0, ""), (74, 100, 0, ""), (74, 124, 0, ""), (74, 148, 0, ""), (74, 74, -1, ""),
            (74, 148, -1, ""), (74, 74, 0, "2:74,2:148"), (74, 74, 0, "1:74,3:148"), (74, 74, 0, "3:74,1:148"),
            (74, 74, 0, "2:74,1:110,1:148"), (74, 74, -1, "2:74,2:148"), (60, 60, 0, "2:60,2:148"), (74, 74, 0, "")]
if len(sys.argv) > 1:
    SETTINGS = [tuple(a.split(";")) for a in sys.argv[1:]]
    SETTINGS = [(int(a), int(b), int(c), d) for a, b, c, d in SETTINGS]
for front, fv, prio, split in SETTINGS:
    os.environ["B200_FRONT_SM_CAP"], os.environ["B200_FV_SM_CAP"] = str(front), str(fv)
    os.environ["B200_ENC_PRIORITY"], os.environ["B200_FV_SPLIT"] = str(prio), split
    m._state, m._graphs, m._side = {}, {}, None  # plans bake the caps, the side stream its priority
    torch.cuda.synchronize()
    ms = timeit()
    print(json.dumps({"front_cap": front, "fv_cap": fv, "enc_priority": prio, "fv_split": split, "ms_per_forward": round(ms, 3),
                      "frames_per_s": round(1000.0 * B / ms, 1)}), flush=True)
