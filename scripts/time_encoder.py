"""Dev tool: timing variants of the out-of-scope cuDNN image encoder (EfficientNetV2-S features, B=4 512x384)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from implicit_depth_b200.bd_model import EffNetV2SFeatures
torch.set_grad_enabled(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
def fuse(m):
    for name, child in list(m.named_children()):
        fuse(child)
    if isinstance(m, nn.Sequential):
        keys = list(m._modules.keys())
        for a, b in zip(keys, keys[1:]):
            if isinstance(m._modules[a], nn.Conv2d) and isinstance(m._modules[b], nn.BatchNorm2d):
                m._modules[a] = torch.nn.utils.fusion.fuse_conv_bn_eval(m._modules[a], m._modules[b])
                m._modules[b] = nn.Identity()
    return m
enc = EffNetV2SFeatures().cuda().eval()
x = torch.randn(4, 3, 384, 512, device="cuda")
ref = enc(x)
print("baseline ms", timeit(lambda: enc(x)))
f = fuse(copy.deepcopy(enc))
out = f(x)
print("fused-bn ms", timeit(lambda: f(x)), "max rel diff", max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out, ref)))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    o = f(x)
print("fused-bn graph ms", timeit(g.replay))
fc = copy.deepcopy(f).to(memory_format=torch.channels_last)
xc = x.contiguous(memory_format=torch.channels_last)
out = fc(xc)
print("fused-bn channels_last ms", timeit(lambda: fc(xc)), "max rel diff", max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out, ref)))
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    o = fc(xc)
print("fused-bn channels_last graph ms", timeit(g2.replay))
torch.backends.cudnn.allow_tf32 = False
print("fused-bn strict-fp32 ms", timeit(lambda: f(x)))
print("fused-bn channels_last strict-fp32 ms", timeit(lambda: fc(xc)))
