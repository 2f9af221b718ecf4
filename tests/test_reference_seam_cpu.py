"""The drop-in seam of SURVEY 8b on the UNMODIFIED reference classes (imported from baseline/_ref, the copy that travels
to the GPU box): `to_b200(manager)` on the reference's managers, and the state-dict contract between `B200BDModel` and
the reference's `BDModel`.  CPU only -- no kernels are launched here."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_loader  # noqa: E402

if not ref_loader.available():
    import subprocess

    subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "make_ref.py")], check=False)
pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not present (no /root/reference here)")


def test_to_b200_on_reference_managers():
    """`model.cost_volume = to_b200(model.cost_volume)` (the swap of test_bd.py:80-81): right type, MLP shared by
    reference like the reference's own to_fast() (cost_volume.py:708-715), identical state-dict keys."""
    from implicit_depth_b200 import B200CostVolumeManager, B200FeatureVolumeManager, to_b200

    _, _, ref_cv, _ = ref_loader.reference_modules()
    for K in (7, 2):
        cin = 16 * (1 + K) + (1 + K) + 3 * (1 + K) + K + K + K + 3 * K
        fv = ref_cv.FeatureVolumeManager(24, 32, num_depth_bins=8, mlp_channels=[cin, 128, 128, 1], matching_dim_size=16,
                                         num_source_views=K)
        for mgr in (fv, fv.to_fast()):
            out = to_b200(mgr)
            assert type(out) is B200FeatureVolumeManager and out.num_source_views == K
            assert out.mlp is mgr.mlp  # moved by reference, not copied
            assert (out.matching_height, out.matching_width, out.num_depth_bins) == (24, 32, 8)
            assert set(out.state_dict()) == set(mgr.state_dict())
            for k, v in mgr.state_dict().items():
                assert out.state_dict()[k].shape == v.shape, k
            assert to_b200(out) is out and out.to_fast() is out
    dot = ref_cv.CostVolumeManager(24, 32, num_depth_bins=8)
    for mgr in (dot, dot.to_fast()):
        out = to_b200(mgr)
        assert type(out) is B200CostVolumeManager and not hasattr(out, "mlp")
        assert set(out.state_dict()) == set(mgr.state_dict())
    # CPU tensors must raise (no fallback), exactly like every other entry point
    from implicit_depth_b200._abi import B200Error

    with pytest.raises(B200Error):
        to_b200(dot)(torch.zeros(1, 16, 24, 32), torch.zeros(1, 2, 16, 24, 32), torch.eye(4).expand(1, 2, 4, 4),
                     torch.eye(4).expand(1, 2, 4, 4), torch.eye(4).expand(1, 2, 4, 4), torch.eye(4)[None],
                     torch.tensor(0.25).view(1, 1, 1, 1), torch.tensor(5.0).view(1, 1, 1, 1))


@pytest.mark.parametrize("kw", [dict(), dict(feature_volume_type="simple_cost_volume"), dict(depth_decoder_name="skip")])
def test_state_dict_is_interchangeable_with_the_reference_bdmodel(kw):
    """`B200BDModel(opts).load_state_dict(reference.state_dict())` and the other way round, strict."""
    from implicit_depth_b200 import synthetic
    from implicit_depth_b200.bd_model import B200BDModel, default_options

    mine = B200BDModel(default_options(image_width=128, image_height=96, matching_num_depth_bins=8, **kw))
    synthetic.init_model_weights(mine, seed=0)
    ref = ref_loader.build_bd_model(128, 96, 8, state_dict=mine.state_dict(), **kw)  # strict load inside
    other = B200BDModel(default_options(image_width=128, image_height=96, matching_num_depth_bins=8, **kw))
    missing, unexpected = other.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    # the seam on a whole reference model: attribute swap keeps the model's key set
    from implicit_depth_b200 import to_b200

    keys = set(ref.state_dict())
    ref.cost_volume = to_b200(ref.cost_volume)
    assert set(ref.state_dict()) == keys
