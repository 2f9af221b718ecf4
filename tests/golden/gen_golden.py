"""Generate golden vectors by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz

Inputs come from `implicit_depth_b200.synthetic` (numpy PCG64, portable), so the fixtures
only store reference OUTPUTS (+ MLP weights, which depend on torch's RNG).  fp32 outputs of
the reference's slow (loop) managers are the goldens; their fast/efficient twins and an
fp64 run are stored as cross-checks / arbiters.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from implicit_depth_b200 import synthetic  # noqa: E402
from modules import cost_volume as ref_cv  # noqa: E402  (reference)

torch.set_grad_enabled(False)

# name: (seed, B, K, C, h, w, D)
VOLUME_CASES = {
    "cfg1_48x64_k2_d16": (1000, 1, 2, 16, 48, 64, 16),
    "small_24x32_k7_d8": (1001, 2, 7, 16, 24, 32, 8),
    "ragged_20x36_k3_d5": (1002, 1, 3, 16, 20, 36, 5),
    "cfg2_frame_96x128_k7_d64": (2000, 1, 7, 16, 96, 128, 64),
}


def t(x, dtype=torch.float32):
    return torch.from_numpy(np.asarray(x)).to(dtype)


def run_manager(mgr, inp, dtype, return_mask):
    mgr = mgr.to(dtype)
    args = {k: t(v, dtype) for k, v in inp.items()}
    mn = torch.tensor(0.25, dtype=dtype).view(1, 1, 1, 1)
    mx = torch.tensor(5.0, dtype=dtype).view(1, 1, 1, 1)
    cv, lowest, planes, mask = mgr(min_depth=mn, max_depth=mx, return_mask=return_mask, **args)
    return cv, lowest, planes[:, :, 0, 0], mask


def volume_goldens():
    for name, (seed, B, K, C, h, w, D) in VOLUME_CASES.items():
        inp = synthetic.make_volume_inputs(seed, B, K, C, h, w)
        out = {}
        # --- dot-product volume: loop manager (golden) + efficient twin ---
        slow = ref_cv.CostVolumeManager(h, w, num_depth_bins=D)
        cv, lowest, planes, _ = run_manager(slow, inp, torch.float32, False)
        out["dot_cost"] = cv.numpy()
        out["dot_lowest"] = lowest.numpy()
        out["planes"] = planes[0].numpy()
        fast = slow.to_fast()
        cvf, _, _, _ = run_manager(fast, inp, torch.float32, False)
        out["dot_slow_vs_fast_maxabs"] = np.array((cv - cvf).abs().max().item())
        cv64, _, _, _ = run_manager(ref_cv.CostVolumeManager(h, w, num_depth_bins=D), inp, torch.float64, False)
        out["dot_cost_f64"] = cv64.numpy()
        # --- MLP feature volume: loop manager (golden) + fast twin ---
        torch.manual_seed(0)
        fv = ref_cv.FeatureVolumeManager(h, w, num_depth_bins=D, mlp_channels=[202, 128, 128, 1],
                                         matching_dim_size=C, num_source_views=K)
        for i, li in enumerate((0, 2, 4)):
            out[f"mlp_w{i}"] = fv.mlp.net[li].weight.numpy().copy()
            out[f"mlp_b{i}"] = fv.mlp.net[li].bias.numpy().copy()
        vol, lowest, _, mask = run_manager(fv, inp, torch.float32, True)
        out["fv_vol"] = vol.numpy()
        out["fv_lowest"] = lowest.numpy()
        out["fv_mask"] = mask.numpy()
        if h * w * D <= 96 * 128 * 16:
            ff = fv.to_fast()
            volf, _, _, maskf = run_manager(ff, inp, torch.float32, True)
            out["fv_slow_vs_fast_maxabs"] = np.array((vol - volf).abs().max().item())
            out["fv_mask_fast_equal"] = np.array(bool((mask == maskf).all()))
        vol64, _, _, _ = run_manager(fv, inp, torch.float64, True)  # converts fv to fp64 in place
        out["fv_vol_f64"] = vol64.numpy()
        path = os.path.join(HERE, f"volume_{name}.npz")
        if "cfg2" in name:  # keep the fixture small: fp16-free, but drop the fp64 copies to a strided sample
            out["dot_cost_f64"] = out["dot_cost_f64"][:, :, ::4, ::4].copy()
            out["fv_vol_f64"] = out["fv_vol_f64"][:, :, ::4, ::4].copy()
        np.savez_compressed(path, **out)
        print(name, {k: (v.shape, v.dtype) for k, v in out.items() if v.ndim > 0},
              "dot slow-vs-fast", out["dot_slow_vs_fast_maxabs"],
              "fv slow-vs-fast", out.get("fv_slow_vs_fast_maxabs"))


if __name__ == "__main__":
    volume_goldens()
