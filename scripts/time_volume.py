"""Kernel timings of the volume path at BASELINE cfg2 (dev tool; prints ms per call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import B200CostVolumeManager, B200FeatureVolumeManager, synthetic
from implicit_depth_b200.cost_volume import _pixel_major

B, K, C, h, w, D = 4, 7, 16, 96, 128, int(os.environ.get("D", 64))
t = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_volume_inputs(2000, B, K, C, h, w).items()}
mn = torch.tensor(0.25, device="cuda").view(1, 1, 1, 1); mx = torch.tensor(5.0, device="cuda").view(1, 1, 1, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
pm = lambda layout: (_pixel_major(t["cur_feats"], B, C, h, w, layout),
                     _pixel_major(t["src_feats"].reshape(B * K, C, h, w), B * K, C, h, w, layout))
cur_pm, src_pm = pm(0)
geo = (t["src_extrinsics"], t["src_poses"], t["src_Ks"], t["cur_invK"], mn, mx, None)
for dot_impl in ("band", "gather"):
    dot = B200CostVolumeManager(h, w, num_depth_bins=D, dot_impl=dot_impl).cuda()
    print(f"dot[{dot_impl}] manager (layout + prep + kernel) ms:", timeit(lambda: dot(min_depth=mn, max_depth=mx, **t)))
    print(f"dot[{dot_impl}] pixel-major (prep + kernel) ms:",
          timeit(lambda: dot.forward_pixel_major(cur_pm, src_pm, *geo, False, B, K, h, w)))
cur_pm, src_pm = pm(1)
for impl in ("tc", "simt"):
    if impl == "simt" and os.environ.get("NO_SIMT"): continue
    fv = B200FeatureVolumeManager(h, w, num_depth_bins=D, impl=impl).cuda()
    print(f"fv[{impl}] manager ms:", timeit(lambda: fv(min_depth=mn, max_depth=mx, return_mask=True, **t)))
    print(f"fv[{impl}] pixel-major (prep + kernel + argmax) ms:", timeit(lambda: fv.forward_pixel_major(cur_pm, src_pm, *geo, True, B, K, h, w)))
