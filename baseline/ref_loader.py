"""Import the unmodified reference from `baseline/_ref/` (see make_ref.py) and build its `BDModel`.

Used by `bench.py`'s `gpu_reference` block and by the seam tests (`tests/test_reference_seam_*.py`).  Nothing under
`implicit_depth_b200/` imports this."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "reference")
SHIMS = os.path.join(HERE, "_ref", "shims")


def available():
    return os.path.isfile(os.path.join(REF, "experiment_modules", "bd_model.py")) and os.path.isdir(SHIMS)


def _paths():
    if not available():
        raise ImportError("baseline/_ref is missing: run `python baseline/make_ref.py` in the build container")
    for p in (SHIMS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)


def reference_modules():
    """(options, bd_model, cost_volume, networks) modules of the reference."""
    _paths()
    import options as ref_options  # noqa: E402  (reference)
    from experiment_modules import bd_model as ref_bd  # noqa: E402
    from modules import cost_volume as ref_cv  # noqa: E402
    from modules import networks as ref_nets  # noqa: E402

    return ref_options, ref_bd, ref_cv, ref_nets


def build_bd_model(image_width=512, image_height=384, num_depth_bins=64, feature_volume_type="mlp_feature_volume",
                   depth_decoder_name="unet_pp", state_dict=None):
    """The reference's `BDModel` (experiment_modules/bd_model.py:37-141) with `configs/models/implicit_depth.yaml`
    values, optionally loaded (strict) with a state dict -- e.g. that of a seeded `B200BDModel`, whose keys are the
    reference's."""
    ref_options, ref_bd, _, _ = reference_modules()
    ro = ref_options.Options()
    ro.image_width, ro.image_height, ro.matching_num_depth_bins = image_width, image_height, num_depth_bins
    ro.feature_volume_type = feature_volume_type
    ro.depth_decoder_name = depth_decoder_name
    ro.binary_loss_positive_weight = 1.0
    ro.bd_edge_regularision = False
    model = ref_bd.BDModel(ro)
    if state_dict is not None:
        model.load_state_dict(dict(state_dict), strict=True)
    return model.eval()
