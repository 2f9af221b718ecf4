"""ctypes binding of the C-ABI library (`include/b200_planesweep.h`).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a
symbol is absent this module raises at import/first use."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200planesweep.so")

c_f = ctypes.c_void_p  # device pointers travel as plain addresses
c_i = ctypes.c_int
c_ll = ctypes.c_longlong

# name -> argument types (return type is always int, 0 = ok)
SIGNATURES = {
    "b200_volume_prepare": [c_f] * 12 + [c_i] * 5 + [ctypes.c_void_p],
    "b200_feats_to_pixel_major": [c_f, c_f, c_i, c_i, c_i, c_ll, c_ll, c_i, ctypes.c_void_p],
    "b200_volume_argmax": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_cv_dot": [c_f] * 7 + [c_i] * 6 + [ctypes.c_void_p],
    "b200_cv_dot_band": [c_f] * 7 + [c_i] * 6 + [ctypes.c_void_p],
    "b200_fv_mlp_simt": [c_f] * 13 + [c_i] * 7 + [ctypes.c_void_p],
    "b200_fv_mlp_tc": [c_f] * 12 + [c_i] * 7 + [ctypes.c_void_p],
    "b200_fv_tc_wimage_bytes": [c_i],
    "b200_fv_tc_layout": [c_i, ctypes.c_void_p],
    "b200_conv_create": [ctypes.c_void_p, ctypes.c_void_p],
    "b200_conv_run": [ctypes.c_void_p, ctypes.c_void_p],
    "b200_conv_destroy": [ctypes.c_void_p],
    "b200_conv_ntile": [c_i],
    "b200_conv_ntile_for": [ctypes.c_void_p],
    "b200_conv_wimage_bytes": [ctypes.c_void_p, ctypes.c_void_p, c_i, c_i, c_i],
    "b200_conv_uses_halo": [ctypes.c_void_p],
    "b200_f32_to_split": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_ll, c_ll, c_ll, c_ll, ctypes.c_void_p],
    "b200_split_to_nchw": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_upsample2x": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_instance_norm": [c_f] * 7 + [c_i] * 6 + [ctypes.c_float, ctypes.c_float, c_i, ctypes.c_void_p],
    "b200_instance_norm_ws_bytes": [c_i, c_i, ctypes.c_void_p, ctypes.c_void_p],
    "b200_stem_conv7": [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_stem_conv7_tc": [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_maxblurpool": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_dwconv3x3_silu": [c_f] * 6 + [c_i] * 6 + [ctypes.c_void_p],
    "b200_squeeze_excite": [c_f] * 10 + [c_i] * 4 + [ctypes.c_void_p],
    "b200_mbconv_dw_se": [c_f] * 13 + [c_i] * 7 + [ctypes.c_void_p],
    "b200_mbconv_pool_block": [],
    "b200_split_add": [c_f] * 6 + [c_ll, ctypes.c_void_p],
    "b200_stem3x3_s2_silu": [c_f] * 5 + [c_i] * 5 + [c_ll] * 4 + [ctypes.c_void_p],
    "b200_stream_create": [c_i, ctypes.POINTER(ctypes.c_void_p)],
    "b200_stream_destroy": [ctypes.c_void_p],
    "b200_channel_dot_exp": [c_f] * 6 + [c_ll, c_i, ctypes.c_void_p],
    "b200_sigmoid_resize": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, ctypes.c_float, c_i, c_i, ctypes.c_void_p],
    "b200_sample_prior": [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_relative_poses": [c_f] * 6 + [c_i, c_i, ctypes.c_void_p],
    "b200_intrinsics_pyramid": [c_f, c_f, c_f, c_i, c_i, ctypes.c_void_p],
    "b200_binary_mlp_create": [ctypes.c_void_p, ctypes.c_void_p],
    "b200_binary_mlp_planes": [ctypes.c_void_p, c_f, c_i, c_f, c_f, ctypes.c_void_p],
    "b200_binary_mlp_search": [ctypes.c_void_p, c_f, c_i, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_f, c_f, c_i,
                               c_f, c_f, ctypes.c_void_p],
    "b200_binary_mlp_destroy": [ctypes.c_void_p],
}
# dev-probe library (csrc/dev/, `build.build_dev()`): tcgen05 self-test + MMA issue-rate probe
DEV_SIGNATURES = {
    "b200_umma_probe": [c_f, c_f, c_f, c_i, c_i, c_i, ctypes.c_void_p],
    "b200_mma_rate": [ctypes.c_void_p, c_i, c_i, c_i, c_i, ctypes.c_void_p],
}

_lib = None


class B200Error(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} not found: build it with `python -m implicit_depth_b200.build` "
            "(there is no CPU/PyTorch fallback for this path)")
    lib = ctypes.CDLL(LIB_PATH)
    _check_digest(lib)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.b200_last_error.restype = ctypes.c_char_p
    lib.b200_abi_version.restype = ctypes.c_int
    lib.b200_conv_wimage_bytes.restype = ctypes.c_longlong
    lib.b200_fv_tc_layout.argtypes = [c_i, ctypes.c_void_p]
    lib.b200_fv_tc_layout.restype = ctypes.c_int
    _lib = lib
    return lib


def _check_digest(lib):
    """A library built from other sources than the ones next to it must not be called through these signatures."""
    from . import build

    if not os.path.isdir(build.CSRC):
        return  # binary-only deployment: nothing to compare with
    try:
        fn = lib.b200_source_digest
    except AttributeError:
        raise B200Error(f"{LIB_PATH} predates the source-digest check: rebuild with `python -m implicit_depth_b200.build`")
    fn.restype = ctypes.c_char_p
    have, want = fn().decode(), build.source_digest()
    if have != want:
        raise B200Error(f"{LIB_PATH} was built from different sources (library {have[:12]}, csrc {want[:12]}): "
                        "rebuild with `python -m implicit_depth_b200.build`")


_dev = None


def load_dev():
    """The dev-probe library (tests/test_umma_probe_gpu.py, scripts/mma_rate.py)."""
    global _dev
    if _dev is None:
        from . import build

        if not os.path.exists(build.DEV_LIB):
            raise B200Error(f"{build.DEV_LIB} not found: build it with `python -m implicit_depth_b200.build`")
        lib = ctypes.CDLL(build.DEV_LIB)
        for name, argtypes in DEV_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        lib.b200_last_error.restype = ctypes.c_char_p
        _dev = lib
    return _dev


def call_dev(name, *args):
    lib = load_dev()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise B200Error(f"{name} failed ({rc}): {lib.b200_last_error().decode()}")


def ptr(t):
    """Device address of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def new_stream(device=None, priority=0):
    """A dedicated CUDA stream (`b200_stream_create`) as a `torch.cuda.ExternalStream`.  Use this, not
    `torch.cuda.Stream()`, for streams that must be distinct inside one CUDA-graph capture: torch hands its streams out
    of a pool of 32 per device and aliases them once a process has created more.  The stream is never destroyed (like
    torch's pooled streams): the caching allocator may still hold blocks keyed by it, and the handful of streams a
    model / pipeline creates live as long as the process anyway."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    load()
    out = ctypes.c_void_p()
    with torch.cuda.device(dev):
        call("b200_stream_create", int(priority), ctypes.byref(out))
    return torch.cuda.ExternalStream(out.value, device=dev)


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise B200Error(f"{name} failed ({rc}): {lib.b200_last_error().decode()}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B200Error("implicit_depth_b200 kernels need CUDA tensors; there is no CPU fallback")
