"""Dev tool: per-op timing of the matching-encoder plan at cfg2 (32 images of 384x512)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200.networks import Plan, ResnetMatchingEncoder
torch.set_grad_enabled(False)
enc = ResnetMatchingEncoder(18, 16).cuda().eval()
n, H, W = 32, 384, 512
img = torch.randn(n, 3, H, W, device="cuda")
g = Plan("cuda"); enc.plan(g, lambda: img, n, H, W)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): g.run()
tot = 0
for i, (op, _, _) in enumerate(g.ops):
    ts = []
    for _ in range(5):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); op(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); tot += ts[2]
    print(f"op {i:2d}: {ts[2]*1e3:8.1f} us")
print(f"sum {tot*1e3:.1f} us")
