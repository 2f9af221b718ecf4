"""Timing of the out-of-scope cuDNN image encoder (EfficientNetV2-S features) in several memory formats (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200.bd_model import EffNetV2SFeatures, fold_batchnorm

torch.set_grad_enabled(False)
enc = EffNetV2SFeatures().cuda().eval()
x = torch.randn(4, 3, 384, 512, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]

def graphed(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): out = fn()
    return g.replay, out

ref = enc(x)
for name, bench, cl, tf32 in [("folded nchw", False, False, True), ("folded nchw benchmark", True, False, True),
                              ("folded channels_last", False, True, True), ("folded channels_last benchmark", True, True, True),
                              ("folded channels_last benchmark fp32", True, True, False)]:
    torch.backends.cudnn.benchmark = bench
    torch.backends.cudnn.allow_tf32 = tf32
    m = fold_batchnorm(enc)
    xi = x
    if cl:
        m = m.to(memory_format=torch.channels_last); xi = x.contiguous(memory_format=torch.channels_last)
    fn = lambda: m(xi)
    out = fn()
    err = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out, ref))
    eager = timeit(fn)
    rep, _ = graphed(fn)
    print(f"{name:40s} eager {eager:7.3f} ms   graph {timeit(rep):7.3f} ms   max rel err vs unfolded {err:.2e}")
