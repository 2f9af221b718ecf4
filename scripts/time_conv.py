"""Timing of representative conv launches (dev tool). B200_CONV_NO_HALO=1 selects the plain kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200.conv import ConvPlan, SplitAct
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(B, H, W, segC, Cout, k=3, hint=0):
    acts = [SplitAct.from_nchw_torch(torch.randn(B, C, H, W, device="cuda")) for C in segC]
    ws = [torch.randn(Cout, C, k, k, device="cuda") * 0.05 for C in segC]
    out = SplitAct(B, H, W, Cout, "cuda")
    plan = ConvPlan([(a, k, 1, k // 2) for a in acts], ws, torch.zeros(Cout, device="cuda"), out, B, Cout, act="lrelu",
                    tile_hint=hint)
    for _ in range(3): plan.run()
    ts = []
    for _ in range(10):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); plan.run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * B * H * W * sum(segC) * k * k * Cout
    print(f"B{B} {H}x{W} {segC}->{Cout} k{k} hint{hint}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s algorithmic")
for hint in (0, 1, 2):  # 64-channel layers: automatic / M=128 items on two CTAs per SM / M=256 items on one
    run(4, 192, 256, [64, 64, 64], 64, hint=hint)
    run(4, 192, 256, [64], 64, hint=hint)
    run(4, 192, 256, [24], 64, hint=hint)
    run(4, 96, 128, [64, 64, 64], 64, hint=hint)
    run(4, 96, 128, [64, 48], 64, hint=hint)
    run(32, 96, 128, [64], 64, hint=hint)
run(4, 48, 64, [128, 128, 128], 128)
run(4, 48, 64, [128], 128)
run(4, 24, 32, [256, 256], 256)
run(4, 12, 16, [384, 256], 384)
run(4, 192, 256, [64], 128, k=1)
