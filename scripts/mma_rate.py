"""Dev probe: cycles per tcgen05.mma (M=128) for different N / operand sources (csrc/mma_rate.cu)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import _abi
lib = _abi.load_dev()
names = {0: "SS split (hi/lo pattern)", 1: "TS split (A in TMEM)", 2: "SS plain, 3 A chunks", 3: "SS same slice",
         4: "SS split, 2 accumulators", 5: "TS split, 2 accumulators", 8: "SW64 halo: merged N + lo N/2 (avg)",
         16: "SW128 merged N + lo N/2 (avg)"}
for grid in (148,):
    for N in (64, 128, 256):
        for mode in (0, 3, 8, 16):
            if N == 256 and (mode & 4):
                continue
            out = torch.zeros(grid, dtype=torch.int64, device="cuda")
            iters = 400
            for _ in range(2):
                rc = lib.b200_mma_rate(out.data_ptr(), N, mode, iters, grid, None)
                assert rc == 0, lib.b200_last_error()
                torch.cuda.synchronize()
            cyc = out.float().mean().item() / (iters * 12)
            print(f"grid {grid:3d} N {N:3d} {names[mode]:28s}: {cyc:7.1f} cycles/MMA (floor {N // 2})")
