"""Latency of the native image-prior encoder plan alone (cfg2: B=4, 512x384), eager and CUDA-graph replay (dev tool)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import synthetic
from implicit_depth_b200.image_encoder import TfEfficientNetV2SFeatures, plan_efficientnet_v2_s
from implicit_depth_b200.networks import Plan

torch.set_grad_enabled(False)
B, H, W = int(os.environ.get("B", 4)), 384, 512
enc = TfEfficientNetV2SFeatures().eval()
synthetic.init_model_weights(enc, seed=3)
img = torch.randn(B, 3, H, W, device="cuda")
g = Plan("cuda")
plan_efficientnet_v2_s(g, enc, lambda: img, B, H, W)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]


eager = timeit(g.run)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    g.run()
torch.cuda.current_stream().wait_stream(s)
cg = torch.cuda.CUDAGraph()
with torch.cuda.graph(cg):
    g.run()
print(json.dumps({"launches": g.n_launches,
                  "eager_ms": round(eager, 3), "graph_ms": round(timeit(cg.replay), 3)}))
