"""Shim (import-time only)."""


class _Stub:
    def __init__(self, *a, **k):
        raise NotImplementedError("pytorch3d shim")


FoVPerspectiveCameras = HardFlatShader = MeshRasterizer = MeshRenderer = RasterizationSettings = _Stub
TexturesAtlas = TexturesVertex = _Stub
