"""gpu-nbody_b200: the B200-native Barnes-Hut step behind gpu-nbody's kernel contract.

Only the hot path lives here: csrc/ (hand-written sm_100a kernels + the C ABI of
include/bhstep.h, built in-tree as libbhstep.so), the ctypes binding (_lib), the
host-side mirror of the reference's simulation classes (simulation), the
universe generators / .universe format (universe) and the multi-GPU driver
(distributed).  No CPU fallback and no dependency on oracle/.
"""
from .simulation import AbstractNBodySimulation, BhError, GPUBarnesHutNBodySimulation, Mode  # noqa: F401
from . import universe  # noqa: F401

__all__ = ["AbstractNBodySimulation", "GPUBarnesHutNBodySimulation", "Mode", "BhError", "universe"]
