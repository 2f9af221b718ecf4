"""Shim (import-time only)."""


class Meshes:
    def __init__(self, *a, **k):
        raise NotImplementedError("pytorch3d shim")
