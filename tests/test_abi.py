"""CPU-side checks of the drop-in boundary: the library builds for sm_100a without a GPU, loads,
and exports every symbol that include/*.h declares; host-side argument errors are reported, not
aborted; the product path refuses to run without CUDA."""
import ctypes
import glob
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(b200_\w+)\s*\(", src))
    return sorted(names)


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "b200_cv_dot" in syms and "b200_volume_prepare" in syms and len(syms) >= 6


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_python_binding_covers_header(lib):
    from implicit_depth_b200 import _abi

    bound = set(_abi.SIGNATURES) | {"b200_last_error", "b200_abi_version", "b200_source_digest"}
    assert set(declared_symbols()) <= bound
    assert lib.b200_abi_version() == 2


def test_library_carries_the_digest_of_its_sources(lib, tmp_path, monkeypatch):
    """A library built from other sources than the csrc/ next to it is refused (no stale binary behind new ctypes
    signatures), and the product library holds no process-global tuning state or dev probes."""
    from implicit_depth_b200 import _abi, build

    lib.b200_source_digest.restype = ctypes.c_char_p
    assert lib.b200_source_digest().decode() == build.source_digest() == build.library_digest()
    monkeypatch.setattr(build, "source_digest", lambda sub="": "0" * 64)
    with pytest.raises(_abi.B200Error, match="different sources"):
        _abi._check_digest(lib)
    for gone in ("b200_set_sm_cap", "b200_sm_cap", "b200_umma_probe", "b200_mma_rate", "b200_conv_set_prof"):
        assert not hasattr(lib, gone), gone
    srcs = "".join(open(f).read() for f in build.sources())
    assert "getenv" not in srcs


def test_to_b200_and_packed_outputs_need_no_gpu():
    """`to_b200` on anything with the reference managers' attributes (the real reference classes are exercised in
    tests/test_reference_seam_cpu.py)."""
    from types import SimpleNamespace

    from implicit_depth_b200 import B200CostVolumeManager, B200FeatureVolumeManager, to_b200
    from implicit_depth_b200.cost_volume import MLP

    dot = to_b200(SimpleNamespace(matching_height=6, matching_width=8, num_depth_bins=4, buffers=lambda: iter(())))
    assert type(dot) is B200CostVolumeManager
    mlp = MLP([26 * 3 + 20, 128, 128, 1], disable_final_activation=True)
    fv = to_b200(SimpleNamespace(matching_height=6, matching_width=8, num_depth_bins=4, mlp=mlp, buffers=lambda: iter(())))
    assert type(fv) is B200FeatureVolumeManager and fv.num_source_views == 3 and fv.mlp is mlp


def test_bad_arguments_return_error_codes(lib):
    rc = lib.b200_cv_dot(None, None, None, None, None, None, None, 1, 7, 8, 4, 4, 4, None)
    assert rc == -1 and b"16 feature channels" in lib.b200_last_error()
    rc = lib.b200_cv_dot(None, None, None, None, None, None, None, 1, 9, 16, 4, 4, 4, None)
    assert rc == -1 and b"bad sizes" in lib.b200_last_error()
    rc = lib.b200_volume_argmax(None, None, None, None, 1, 1, 1, None)
    assert rc == -1


def test_sass_contains_tcgen05(lib):
    """The tensor-core kernels must really be tcgen05 (UTC*MMA / LDTM / STTM in SASS)."""
    import shutil
    import subprocess

    from implicit_depth_b200 import _abi

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "STTM" in sass and "LDTM" in sass


def test_product_refuses_cpu_tensors():
    from implicit_depth_b200 import B200CostVolumeManager, _abi

    m = B200CostVolumeManager(4, 4, num_depth_bins=2)
    z = torch.zeros
    with pytest.raises(_abi.B200Error):
        m(z(1, 16, 4, 4), z(1, 1, 16, 4, 4), z(1, 1, 4, 4), z(1, 1, 4, 4), z(1, 1, 4, 4), z(1, 4, 4),
          torch.tensor(0.25).view(1, 1, 1, 1), torch.tensor(5.0).view(1, 1, 1, 1))


def test_state_dict_keys_match_reference():
    """Keys recorded from the reference managers (checkpoint row)."""
    from implicit_depth_b200 import B200CostVolumeManager, B200FeatureVolumeManager

    fv = B200FeatureVolumeManager(6, 8, num_depth_bins=4)
    keys = set(fv.state_dict().keys())
    want = {"linear_ramp_1d11", "backprojector.pix_coords_13N", "projector.eps"} | {
        f"mlp.net.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias")}
    assert keys == want
    assert tuple(fv.mlp.net[0].weight.shape) == (128, 202)
    assert set(B200CostVolumeManager(6, 8).state_dict().keys()) == {
        "linear_ramp_1d11", "backprojector.pix_coords_13N", "projector.eps"}


def test_channel_permutation_is_a_permutation_of_the_variable_channels():
    from implicit_depth_b200 import B200FeatureVolumeManager

    for K in (1, 2, 3, 7):
        perm = B200FeatureVolumeManager.channel_permutation(K)
        cin = 26 * K + 20
        const = set(range(16 * (K + 1), 16 * (K + 1) + K)) | set(range(cin - 3 * K, cin))
        assert len(perm) == 22 * K + 20 == len(set(perm))
        assert set(perm) | const == set(range(cin)) and not (set(perm) & const)


def test_tc_channel_layout_covers_every_variable_channel_once():
    """K-dimension layout of the tensor-core feature-volume kernel: every (pixel, plane)-dependent reference channel
    appears exactly once, padding only at the end of each role's share, whole 64-channel chunks."""
    from implicit_depth_b200 import B200FeatureVolumeManager

    for K in range(1, 9):
        ref = B200FeatureVolumeManager.channel_permutation(K)
        lay = B200FeatureVolumeManager.tc_channel_layout(K)
        real = [c for c in lay if c >= 0]
        assert sorted(real) == sorted(ref) and len(lay) % 64 == 0
        assert _abi_bytes(K) == (2 * (len(lay) // 64) + 4) * 16384


def _abi_bytes(K):
    from implicit_depth_b200 import _abi

    return _abi.load().b200_fv_tc_wimage_bytes(K)


def test_product_never_touches_the_oracle():
    """`oracle/` is test infrastructure: nothing under the product package may import or execute it, and the only
    places outside tests/ that do are smoke() and bench.py's CPU-baseline legs."""
    pkg = os.path.join(ROOT, "implicit_depth_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle." not in src and "/oracle" not in src, f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"^\s*(from|import)\s+oracle\b", bench, flags=re.M)]
    assert len(uses) == 1 and bench[:uses[0]].rfind("def cpu_reference_run") > bench[:uses[0]].rfind("def main_b200")
