"""Shim (import-time only)."""


def load_ply(*a, **k):
    raise NotImplementedError("pytorch3d shim")
