"""One eager (no CUDA graph) forward step at BASELINE cfg2 between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options

torch.set_grad_enabled(False)
B = int(os.environ.get("B", 4))
opts = default_options(image_width=512, image_height=384, matching_num_depth_bins=64)
model = B200BDModel(opts)
synthetic.init_model_weights(model, seed=0)
model = model.cuda().eval()
cur, src = synthetic.make_frame_batch(2000, B, 7, 384, 512)
# staged inputs (implicit_depth_b200.staging), like bench.py: the forward reads the device slot in place
from implicit_depth_b200.staging import FrameStaging

staging = FrameStaging(B, 7, 384, 512, P=8)
frame = staging.device_frame("cuda")
FrameStaging.upload(staging.host_frame().fill(cur, src), frame)
cur, src = frame.cur, frame.src
for _ in range(2):
    model("test", cur, src, return_mask=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model("test", cur, src, return_mask=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
