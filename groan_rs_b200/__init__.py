"""groan_rs_b200 -- B200-native (sm_100a) implementation of groan_rs's per-frame PBC geometry hot path.

The product is libgroan_gpu.so (hand-written CUDA + the C ABI of include/groan_gpu.h); this package is the
thin Python host side that mirrors the reference's System / Group / SimBox / Dimension surface for that
path.  Importing the package does not load the library; creating a System does, and fails loudly when the
CUDA library is missing (there is no CPU fallback).
"""
from ._lib import FLAG_EXACT_ONLY, FLAG_HOST_FALLBACK, FLAG_NO_TMA, FLAG_TRICLINIC, GroanLibraryMissing  # noqa: F401
from .system import (Dimension, GpuError, GroanError, Group, GroupError, MassError, PositionError, RMSDError,  # noqa: F401
                     SimBox, SimBoxError, System)
from . import xtc  # noqa: F401
from .parallel import frame_range, gather_frames, traj_iter_map_reduce  # noqa: F401

__all__ = ["System", "Group", "SimBox", "Dimension", "GroanError", "GroupError", "SimBoxError", "PositionError",
           "MassError", "RMSDError", "GpuError", "frame_range", "gather_frames", "traj_iter_map_reduce",
           "xtc", "FLAG_TRICLINIC", "FLAG_EXACT_ONLY", "FLAG_NO_TMA", "FLAG_HOST_FALLBACK", "GroanLibraryMissing"]
