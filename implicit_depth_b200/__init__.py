"""implicit_depth_b200 -- B200-native (sm_100a) plane-sweep hot path for implicit-depth.

Product code only: CUDA kernels behind a C ABI (`csrc/`, `include/b200_planesweep.h`) and the
host-side mirror of the reference's module interface.  No CPU / PyTorch fallback."""
from .cost_volume import (  # noqa: F401
    B200CostVolumeManager,
    B200FeatureVolumeManager,
    to_b200,
)
from . import _abi  # noqa: F401

__all__ = ["B200CostVolumeManager", "B200FeatureVolumeManager", "to_b200"]
