"""Dev tool: per-role clock64 breakdown of conv_halo_kernel (b200_conv_set_prof)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import _abi
from implicit_depth_b200.conv import ConvPlan, SplitAct
lib = _abi.load()
lib.b200_conv_set_prof.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
NAMES = ["prod:wait p_empty", "prod:wait b_empty", "mma:wait acc_empty", "mma:wait p_full", "mma:wait b_full",
         "epi:wait acc_full", "epi:body", "total(mma warp)"]
def run(B, H, W, segC, Cout, k=3):
    acts = [SplitAct.from_nchw_torch(torch.randn(B, C, H, W, device="cuda")) for C in segC]
    ws = [torch.randn(Cout, C, k, k, device="cuda") * 0.05 for C in segC]
    out = SplitAct(B, H, W, Cout, "cuda")
    plan = ConvPlan([(a, k, 1, k // 2) for a in acts], ws, torch.zeros(Cout, device="cuda"), out, B, Cout, act="lrelu")
    buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    grid = lib.b200_conv_set_prof(plan.handle, buf.data_ptr())
    for _ in range(3): plan.run()
    torch.cuda.synchronize()
    r = buf.view(148, 8)[:grid].double()
    print(f"B{B} {H}x{W} {segC}->{Cout} k{k} grid {grid}")
    for i, n in enumerate(NAMES):
        print(f"   {n:22s} mean {r[:, i].mean().item():10.0f}  max {r[:, i].max().item():10.0f}")
run(4, 192, 256, [64, 64, 64], 64)
run(4, 192, 256, [64], 64)
run(4, 48, 64, [128, 128, 128], 128)
run(4, 192, 256, [64], 128, k=1)
