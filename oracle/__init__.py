"""ctypes wrapper around oracle/bh_oracle.c -- TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from gpu_nbody_b200/.
The buffers are numpy arrays in the reference's own SoA layout
(GPUBarnesHutNBodySimulation.java:153-181), so tests compare them directly
with what the CUDA library's bh_read returns.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BUFFER_NAMES = ["posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ",
                "step", "blockCount", "bodyCount", "radius", "maxDepth", "bottom", "mass",
                "child", "start", "sorted", "error"]
_FLOAT = {"posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "radius", "mass"}
_SCALAR = {"step", "blockCount", "radius", "maxDepth", "bottom", "error"}


class _State(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in BUFFER_NAMES] +
                [("n", C.c_int32), ("m", C.c_int32), ("theta_macro", C.c_float), ("epsilon", C.c_float),
                 ("timestep", C.c_float), ("vote_width", C.c_int32), ("fma_policy", C.c_int32),
                 ("interactions", C.c_int64), ("opens", C.c_int64), ("group_interactions", C.c_void_p)])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libbh_oracle.so")
    src = os.path.join(_HERE, "bh_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libbh_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.bho_number_of_nodes.restype = C.c_int32
        _LIB.bho_number_of_nodes.argtypes = [C.c_int32]
        for name in ("bho_bounding_box", "bho_summarize", "bho_sort", "bho_integrate"):
            getattr(_LIB, name).restype = None
            getattr(_LIB, name).argtypes = [C.POINTER(_State)]
        for name in ("bho_build_tree", "bho_build_tree_parallel", "bho_calculate_force"):
            getattr(_LIB, name).restype = C.c_int32
            getattr(_LIB, name).argtypes = [C.POINTER(_State)]
        _LIB.bho_calculate_force_range.restype = C.c_int32
        _LIB.bho_calculate_force_range.argtypes = [C.POINTER(_State), C.c_int32, C.c_int32]
        _LIB.bho_step.restype = C.c_int32
        _LIB.bho_step.argtypes = [C.POINTER(_State), C.c_int32]
        _LIB.bho_canonicalize.restype = C.c_int32
        _LIB.bho_canonicalize.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _LIB.bho_energy.restype = None
        _LIB.bho_energy.argtypes = [C.c_int32] + [C.c_void_p] * 7 + [C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _LIB.bho_direct_acc.restype = None
        _LIB.bho_direct_acc.argtypes = [C.c_int32] + [C.c_void_p] * 4 + [C.c_float, C.c_int32, C.c_int32] + [C.c_void_p] * 3
        _LIB.bho_num_threads.restype = C.c_int32
        _LIB.bho_set_num_threads.restype = None
        _LIB.bho_set_num_threads.argtypes = [C.c_int32]
    return _LIB


def number_of_nodes(nbodies: int) -> int:
    return int(lib().bho_number_of_nodes(nbodies))


def num_threads() -> int:
    return int(lib().bho_num_threads())


def use_all_cores() -> int:
    """OpenMP threads = the cores this process may run on (torchrun sets OMP_NUM_THREADS=1 for its workers)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().bho_set_num_threads(n)
    return num_threads()


class OracleSim:
    """The reference's buffer set + its six kernels, on the CPU.

    theta is the opening angle; the reference's THETA macro is theta**2
    (calculateforce.cl:15-16, SURVEY.md 7 hard part 2).
    """

    def __init__(self, nbodies, x, y, z, vx, vy, vz, mass, theta=0.5, eps2=0.0025, dt=0.025,
                 vote_width=16, fma_policy=1, theta_macro=None):
        self.lib = lib()
        self.n = int(nbodies)
        self.m = number_of_nodes(self.n)
        m1 = self.m + 1
        self.buf = {}
        for name in BUFFER_NAMES:
            dt_ = np.float32 if name in _FLOAT else np.int32
            size = 1 if name in _SCALAR else (8 * m1 if name == "child" else m1)
            self.buf[name] = np.zeros(size, dtype=dt_)
        self.buf["step"][0] = -1          # GPUBH:165
        self.buf["maxDepth"][0] = 1       # GPUBH:170
        for name, src in zip(("posX", "posY", "posZ", "velX", "velY", "velZ", "mass"), (x, y, z, vx, vy, vz, mass)):
            self.buf[name][: self.n] = np.asarray(src, dtype=np.float32)[: self.n]
        self.group_interactions = np.zeros((self.n + vote_width - 1) // vote_width, dtype=np.int32)
        self.state = _State()
        for name in BUFFER_NAMES:
            setattr(self.state, name, self.buf[name].ctypes.data)
        self.state.n, self.state.m = self.n, self.m
        self.state.theta_macro = float(np.float32(theta) * np.float32(theta)) if theta_macro is None else theta_macro
        self.state.epsilon, self.state.timestep = eps2, dt
        self.state.vote_width, self.state.fma_policy = vote_width, fma_policy
        self.state.group_interactions = self.group_interactions.ctypes.data

    def __getattr__(self, name):
        buf = self.__dict__.get("buf", {})
        if name in buf:
            return buf[name]
        raise AttributeError(name)

    # the six kernels, GPUBH:258-263
    def bounding_box(self): self.lib.bho_bounding_box(C.byref(self.state))
    def build_tree(self): return self.lib.bho_build_tree(C.byref(self.state))
    def build_tree_parallel(self): return self.lib.bho_build_tree_parallel(C.byref(self.state))
    def summarize(self): self.lib.bho_summarize(C.byref(self.state))
    def sort(self): self.lib.bho_sort(C.byref(self.state))
    def calculate_force(self): return self.lib.bho_calculate_force(C.byref(self.state))
    def calculate_force_range(self, first, count): return self.lib.bho_calculate_force_range(C.byref(self.state), first, count)
    def integrate(self): self.lib.bho_integrate(C.byref(self.state))
    def step(self, nsteps=1): return self.lib.bho_step(C.byref(self.state), nsteps)

    @property
    def interactions(self): return int(self.state.interactions)

    @property
    def opens(self): return int(self.state.opens)

    @property
    def cells_used(self): return self.m - int(self.buf["bottom"][0]) + 1

    def canonical(self):
        return canonicalize(self.buf["child"], self.n, self.m)

    def energy(self):
        return energy(*(self.buf[k][: self.n] for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "mass")),
                      eps2=float(self.state.epsilon))


def canonicalize(child, n, m, max_cells=None):
    """Returns (order, canon): order[c] = original index of the c-th cell in DFS
    pre-order from the root, canon[c, k] = -1 | body | n + DFS number of child cell."""
    child = np.ascontiguousarray(child, dtype=np.int32)
    max_cells = (m - n + 1) if max_cells is None else max_cells
    order = np.empty(max_cells, dtype=np.int32)
    canon = np.empty(8 * max_cells, dtype=np.int32)
    c = lib().bho_canonicalize(child.ctypes.data, n, m, max_cells, order.ctypes.data, canon.ctypes.data)
    if c < 0:
        raise ValueError("child array is not a tree rooted at %d" % m)
    return order[:c].copy(), canon[: 8 * c].reshape(c, 8).copy()


def energy(x, y, z, vx, vy, vz, mass, eps2=0.0025):
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z, vx, vy, vz, mass)]
    ek, ep = C.c_double(), C.c_double()
    lib().bho_energy(arrs[0].size, *(a.ctypes.data for a in arrs), eps2, C.byref(ek), C.byref(ep))
    return ek.value, ep.value


def direct_acc(x, y, z, mass, first, count, eps2=0.0025):
    arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z, mass)]
    out = [np.empty(count, dtype=np.float64) for _ in range(3)]
    lib().bho_direct_acc(arrs[0].size, *(a.ctypes.data for a in arrs), eps2, first, count, *(o.ctypes.data for o in out))
    return out
