"""The SURVEY 8b seam exercised end to end on the GPU with the UNMODIFIED reference (baseline/_ref, shipped by
baseline/make_ref.py): the reference's own `BDModel` (PyTorch/cuDNN, strict fp32) run as `test_bd.py` runs it, then
with `model.cost_volume = to_b200(model.cost_volume)` swapped in (test_bd.py:80-81), then `B200BDModel` with the same
state dict -- all three must agree within 1e-3, plane indices exact outside fp64 near-ties."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_loader  # noqa: E402

from implicit_depth_b200 import synthetic, to_b200  # noqa: E402
from implicit_depth_b200.bd_model import B200BDModel, default_options  # noqa: E402

from cases import argmax_exactness, mask_exactness, record, rel_err  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not shipped with this snapshot")]
TOL = 1e-3


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("fv_type,K,fast", [("mlp_feature_volume", 7, False), ("mlp_feature_volume", 7, True),
                                            ("simple_cost_volume", 3, False)])
def test_reference_bdmodel_with_b200_volume_swapped_in(fv_type, K, fast):
    H, W, D, B = 192, 256, 16, 2
    opts = default_options(image_width=W, image_height=H, matching_num_depth_bins=D, feature_volume_type=fv_type,
                           num_source_views=K)
    mine = B200BDModel(opts)
    synthetic.init_model_weights(mine, seed=0)
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    ref = ref_loader.build_bd_model(W, H, D, feature_volume_type=fv_type, state_dict=sd)
    if fast:
        ref.cost_volume = ref.cost_volume.to_fast()  # test_bd.py:80-81 (--fast_cost_volume)
    ref = ref.cuda().eval()
    cur, src = synthetic.make_frame_batch(8100 + K, B, K, H, W)
    c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    s = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    seen = {}
    hook = ref.cost_volume.register_forward_pre_hook(lambda m, a, kw: seen.update(kw), with_kwargs=True)
    with torch.inference_mode():
        want = ref("test", dict(c), s, unbatched_matching_encoder_forward=not fast, return_mask=True)
    hook.remove()
    # fp64 arbiter: the reference's own manager in double on the features its encoder produced
    arb_mgr = ref_loader.reference_modules()[2]
    if fv_type == "mlp_feature_volume":
        a = arb_mgr.FeatureVolumeManager(H // 4, W // 4, num_depth_bins=D, mlp_channels=[202, 128, 128, 1],
                                         matching_dim_size=16, num_source_views=K)
        a.mlp.load_state_dict(ref.cost_volume.mlp.state_dict())
    else:
        a = arb_mgr.CostVolumeManager(H // 4, W // 4, num_depth_bins=D)
    a = a.cuda().double()
    with torch.inference_mode():
        arb = a(**{k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in seen.items()})[0]
    arb = arb.cpu().numpy()
    # ---- the seam: swap the manager on the reference model, nothing else changes ----
    ref.cost_volume = to_b200(ref.cost_volume)
    assert next(ref.cost_volume.buffers()).is_cuda
    with torch.inference_mode():
        got = ref("test", dict(c), s, unbatched_matching_encoder_forward=not fast, return_mask=True)
    planes = np.exp(np.linspace(np.log(0.25), np.log(5.0), D))
    to_idx = lambda z: np.abs(np.log(z.cpu().numpy())[..., None] - np.log(planes)).argmin(-1)
    case = f"reference_seam/{fv_type}{'_fast' if fast else ''}"
    e = rel_err(got["pred_0"].cpu().numpy(), want["pred_0"].cpu().numpy())
    record(case, kind="pred_0", rel_err=e)
    assert e < TOL
    n_bad, n_near = argmax_exactness(case, to_idx(got["lowest_cost_bhw"]), to_idx(want["lowest_cost_bhw"]), arb)
    assert n_bad == n_near, f"{n_bad} plane-index mismatches, only {n_near} at fp64 near-ties"
    if want["overall_mask_bhw"] is not None:
        from oracle import planesweep_torch as PT  # checker only

        edge = PT.mask_edge_distance(seen["src_extrinsics"].double().cpu(), seen["src_Ks"].double().cpu(),
                                     seen["cur_invK"].double().cpu(), float(planes[-1]), H // 4, W // 4).numpy()
        m_bad, m_edge = mask_exactness(case, got["overall_mask_bhw"].cpu().numpy(),
                                       want["overall_mask_bhw"].cpu().numpy(), edge)
        assert m_bad == m_edge
    else:
        assert got["overall_mask_bhw"] is None
    # ---- the whole-model replacement with the same state dict ----
    if not fast:
        m = mine.cuda().eval()
        out = m("test", dict(c), s, unbatched_matching_encoder_forward=True, return_mask=True)
        e2 = rel_err(out["pred_0"].cpu().numpy(), want["pred_0"].cpu().numpy())
        record(case + "/B200BDModel", kind="pred_0", rel_err=e2)
        assert e2 < TOL
