"""Dev tool: timing of the fused binary-MLP kernel at cfg2 (B=4, 192x256, 8 planes) and of the bisection."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200.conv import SplitAct
from implicit_depth_b200.networks import BinaryMLPNetwork, Plan
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
B, H, W, P = 4, 192, 256, 8
net = BinaryMLPNetwork([64, 64, 128, 256]).cuda()
fa = SplitAct.from_nchw_torch(torch.randn(B, 64, H, W, device="cuda"))
d = torch.rand(B, P, H, W, device="cuda") * 4 + 1
g = Plan("cuda"); pred = net.plan_val(g, fa, lambda: d, P)
ms = timeit(g.run)
fl = 2.0 * B * H * W * P * (65 * 128 + 128 * 128 + 128)
print(f"planes P={P}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s algorithmic")
g2 = Plan("cuda"); s, l = net.plan_search(g2, fa)
ms = timeit(g2.run)
print(f"search 12 iters: {ms*1e3:.1f} us")
