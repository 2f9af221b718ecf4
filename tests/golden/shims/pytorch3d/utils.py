"""Shim (import-time only)."""


def cameras_from_opencv_projection(*a, **k):
    raise NotImplementedError("pytorch3d shim")
