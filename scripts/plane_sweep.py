"""BASELINE config 5: depth-plane sweep 16/32/64/128/256 at 512x384 (96x128 matching), 7 views, B=4 on one B200.
Prints one JSON line per (kernel, D): ms per launch, achieved algorithmic HBM GB/s (cv_dot) / TFLOP/s (fv_tc)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import B200CostVolumeManager, B200FeatureVolumeManager, synthetic

B, K, C, h, w = 4, 7, 16, 96, 128
import importlib.util
_spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bench.py"))
_bench = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(_bench)
_h, _t, _ = _bench.load_peaks()
peaks = {"hbm_gbs": _h, "bf16_tflops": _t}
t = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_volume_inputs(2000, B, K, C, h, w).items()}
mn = torch.tensor(0.25, device="cuda").view(1, 1, 1, 1); mx = torch.tensor(5.0, device="cuda").view(1, 1, 1, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
for D in (16, 32, 64, 128, 256):
    dot = B200CostVolumeManager(h, w, num_depth_bins=D).cuda()
    ms = timeit(lambda: dot(min_depth=mn, max_depth=mx, **t))
    by = 4.0 * h * w * (C * (K + 1) + D) * B
    print(json.dumps({"cfg": 5, "kernel": "cv_dot (manager: layout + prepare + kernel)", "D": D, "ms": ms,
                      "algorithmic_GBps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peaks["hbm_gbs"],
                      "gather_GBps": 4.0 * 16 * 4 * K * D * h * w * B / ms / 1e6}))
    fv = B200FeatureVolumeManager(h, w, num_depth_bins=D).cuda()
    ms = timeit(lambda: fv(min_depth=mn, max_depth=mx, return_mask=True, **t))
    fl = 2.0 * D * h * w * (202 * 128 + 128 * 128 + 128) * B
    print(json.dumps({"cfg": 5, "kernel": "fv_tc (manager: layout + prepare + kernel + argmax)", "D": D, "ms": ms,
                      "algorithmic_TFLOPs": fl / ms / 1e9, "frac_of_bf16_peak": fl / ms / 1e9 / peaks["bf16_tflops"]}))
