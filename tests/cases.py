"""Shared case tables for the volume tests (same seeds as tests/golden/gen_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name: (seed, B, K, C, h, w, D)
VOLUME_CASES = {
    "cfg1_48x64_k2_d16": (1000, 1, 2, 16, 48, 64, 16),
    "small_24x32_k7_d8": (1001, 2, 7, 16, 24, 32, 8),
    "ragged_20x36_k3_d5": (1002, 1, 3, 16, 20, 36, 5),
    "cfg2_frame_96x128_k7_d64": (2000, 1, 7, 16, 96, 128, 64),
}


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"volume_{name}.npz"))


def mlp_weights(g):
    return [(g[f"mlp_w{i}"], g[f"mlp_b{i}"]) for i in range(3)]


def rel_err(a, b):
    """max |a-b| / max |b|: the '1e-3 relative' of BASELINE.md section 6."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def argmax_report(idx, vol_ref64=None, vol_ref=None):
    """Mismatch count between plane indices and the (fp32) reference argmax, and how many of the
    mismatches are near-ties (top-2 margin of the reference below 1e-5 of its range)."""
    ref_idx = np.argmax(vol_ref, axis=1)
    bad = idx != ref_idx
    n_bad = int(bad.sum())
    if n_bad == 0:
        return 0, 0
    srt = np.sort(vol_ref, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    scale = vol_ref.max() - vol_ref.min()
    near = margin[bad] <= 1e-5 * scale
    return n_bad, int(near.sum())
