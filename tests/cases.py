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


# ---- exactness bookkeeping (north_star: "bit-exact for the depth-plane index argmax") -------------------------
# An index mismatch is accepted only at a NEAR-TIE of the fp64 arbiter: the arbiter's values at the two candidate
# planes differ by at most `tie_tol` of the volume's range (the fp32 reference itself flips such pixels against its
# own fp64 run).  Every call records what it saw so that "exact outside fp64 near-ties" is a committed number
# (profiles/r02_parity_exactness.jsonl), not an allowance.
TIE_TOL = 1e-5
_LOG = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out", "parity_exactness.jsonl")


def record(case, **numbers):
    import json

    row = {"case": case}
    row.update({k: (float(v) if isinstance(v, (np.floating, float)) else int(v) if isinstance(v, (np.integer, int))
                    and not isinstance(v, bool) else v) for k, v in numbers.items()})
    print("PARITY", json.dumps(row))
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps(row) + "\n")
    except OSError:
        pass
    return row


def argmax_exactness(case, idx_got, idx_ref, arbiter, tie_tol=TIE_TOL, stride=1):
    """idx_got / idx_ref [B,h,w] plane indices (product / fp32 reference); arbiter [B,D,h',w'] the fp64 volume
    (possibly a ::stride spatial sample of the full map).  Returns (n_bad, n_near): mismatches, and how many of
    them the arbiter calls a tie between the two candidate planes.  Also records the arbiter's smallest top-2
    margin over all pixels (how close to a tie the data gets at all)."""
    ig, ir = idx_got[:, ::stride, ::stride], idx_ref[:, ::stride, ::stride]
    assert ig.shape == arbiter[:, 0].shape, (ig.shape, arbiter.shape)
    scale = float(arbiter.max() - arbiter.min())
    a_got = np.take_along_axis(arbiter, ig[:, None], 1)[:, 0]
    a_ref = np.take_along_axis(arbiter, ir[:, None], 1)[:, 0]
    bad = ig != ir
    gap = np.abs(a_got - a_ref)[bad] / max(scale, 1e-300)
    srt = np.sort(arbiter, axis=1)
    top2 = (srt[:, -1] - srt[:, -2]) / max(scale, 1e-300)
    n_bad, n_near = int(bad.sum()), int((gap <= tie_tol).sum())
    record(case, kind="argmax", pixels=int(bad.size), n_bad=n_bad, n_near=n_near,
           worst_bad_gap_rel=float(gap.max()) if n_bad else 0.0, min_top2_margin_rel=float(top2.min()),
           tie_tol=tie_tol, n_bad_full_map=int((idx_got != idx_ref).sum()))
    return n_bad, n_near


def mask_exactness(case, mask_got, mask_ref, edge_dist=None, edge_tol=1e-3):
    """Boolean-map mismatches, and how many of them sit within `edge_tol` pixels of the decision boundary
    (`edge_dist` [B,h,w]: fp64 distance of the deciding projection to the 2 / w-2 / h-2 window edge)."""
    bad = mask_got != mask_ref
    n_bad = int(bad.sum())
    n_edge = int((edge_dist[bad] <= edge_tol).sum()) if (edge_dist is not None and n_bad) else 0
    record(case, kind="mask", pixels=int(bad.size), n_bad=n_bad, n_at_edge=n_edge,
           worst_bad_edge_dist=float(edge_dist[bad].max()) if (edge_dist is not None and n_bad) else 0.0,
           edge_tol=edge_tol)
    return n_bad, n_edge
