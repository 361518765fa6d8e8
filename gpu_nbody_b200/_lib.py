"""Build and load libbhstep.so (the C ABI of include/bhstep.h) through ctypes.

This is the binding a host language writes against the C ABI; the Java
equivalent (Panama FFM) is java/ch/fhnw/woipv/nbody/simulation/gpu/BhStep.java
and INTEGRATION.md.  There is no fallback: if the library is missing or does
not load, importing callers get an exception.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libbhstep.so")
SOURCES = [os.path.join(PKG_DIR, "csrc", f) for f in ("bhstep.cu", "bh_kernels.cuh")]
HEADER = os.path.join(ROOT, "include", "bhstep.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

BUFFERS = ["posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "step", "blockCount",
           "bodyCount", "radius", "maxDepth", "bottom", "mass", "child", "start", "sorted", "error"]
FLOAT_BUFFERS = {"posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "radius", "mass"}
STAGES = ["bounding_box", "build_tree", "summarize", "sort", "calculate_force", "integrate"]


class BhDiag(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("ekin", "epot", "px", "py", "pz", "mass")]


class BhStats(C.Structure):
    _fields_ = [("nbodies", C.c_int32), ("number_of_nodes", C.c_int32), ("cells_used", C.c_int32),
                ("max_depth", C.c_int32), ("step", C.c_int32), ("error", C.c_int32),
                ("steps_timed", C.c_int64), ("stage_ms", C.c_double * 6), ("stage_launches", C.c_int64 * 6),
                ("interactions", C.c_int64), ("opens", C.c_int64), ("barrier_ms", C.c_double), ("deep_walk", C.c_int32),
                ("walk_spills", C.c_int32)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    deps = SOURCES + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return LIB_PATH
    if not os.path.exists(SOURCES[0]):
        if os.path.exists(LIB_PATH):
            return LIB_PATH
        raise FileNotFoundError(SOURCES[0])
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-o", LIB_PATH, SOURCES[0]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def load():
    """dlopen libbhstep.so and declare every prototype of include/bhstep.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("BHSTEP_LIBRARY") or build()  # BHSTEP_LIBRARY: an alternative build of the same sources (kernel experiments)
    lib = C.CDLL(path)
    p, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    protos = {
        "bh_create": (C.c_int, [C.POINTER(p), i32, f32, f32, f32, i32, i32]),
        "bh_destroy": (None, [p]),
        "bh_last_error": (C.c_char_p, [p]),
        "bh_set_theta_macro": (C.c_int, [p, f32]),
        "bh_set_stream": (C.c_int, [p, p]),
        "bh_use_private_stream": (C.c_int, [p]),
        "bh_set_profiling": (C.c_int, [p, i32]),
        "bh_set_counting": (C.c_int, [p, i32]),
        "bh_set_insertion_order": (C.c_int, [p, i32]),
        "bh_set_graph": (C.c_int, [p, i32]),
        "bh_upload": (C.c_int, [p] + [p] * 7),
        "bh_upload_device": (C.c_int, [p] + [p] * 7),
        "bh_upload_async": (C.c_int, [p] + [p] * 7),
        "bh_bounding_box": (C.c_int, [p]),
        "bh_build_tree": (C.c_int, [p]),
        "bh_summarize": (C.c_int, [p]),
        "bh_sort": (C.c_int, [p]),
        "bh_calculate_force": (C.c_int, [p]),
        "bh_integrate": (C.c_int, [p]),
        "bh_step": (C.c_int, [p, i32]),
        "bh_step_async": (C.c_int, [p, i32]),
        "bh_check": (C.c_int, [p]),
        "bh_calculate_force_slice": (C.c_int, [p, i32, i32]),
        "bh_apply_acceleration": (C.c_int, [p]),
        "bh_finish_async": (C.c_int, [p]),
        "bh_ipc_clear_peers": (C.c_int, [p]),
        "bh_set_slice": (C.c_int, [p, i32, i32]),
        "bh_peer_barrier": (C.c_int, [p]),
        "bh_set_force_deep_walk": (C.c_int, [p, i32]),
        "bh_copy_vertices_device": (C.c_int, [p, p, p]),
        "bh_set_vertex_buffers": (C.c_int, [p, p, p]),
        "bh_write_universe_file": (C.c_int, [p, C.c_char_p]),
        "bh_acc_sorted_device_ptr": (p, [p]),
        "bh_ipc_export": (C.c_int, [p, p]),
        "bh_ipc_set_peers": (C.c_int, [p, i32, i32, p]),
        "bh_calculate_force_slice_p2p": (C.c_int, [p, i32, i32]),
        "bh_stage_async": (C.c_int, [p, i32]),
        "bh_read": (C.c_int, [p, i32, p, i64]),
        "bh_buffer_length": (i64, [p, i32]),
        "bh_copy_vertices": (C.c_int, [p, p, p]),
        "bh_copy_vertices_async": (C.c_int, [p, p, p]),
        "bh_wait_copies": (C.c_int, [p]),
        "bh_stats": (C.c_int, [p, C.POINTER(BhStats)]),
        "bh_diagnostics": (C.c_int, [p, i32, C.POINTER(BhDiag)]),
        "bh_generate_universe": (C.c_int, [p, i32, C.c_uint64, f32, f32, f32]),
        "bh_universe_file_bodies": (C.c_int, [C.c_char_p, C.POINTER(i32)]),
        "bh_upload_universe_file": (C.c_int, [p, C.c_char_p]),
        "bh_read_universe_file": (C.c_int, [C.c_char_p, i32] + [p] * 7),
        "bh_reset_stats": (C.c_int, [p]),
        "bh_number_of_bodies": (i32, [p]),
        "bh_number_of_nodes": (i32, [i32]),
        "bh_abi_version": (i32, []),
        "bh_measure_fp32_peak": (C.c_int, [i32, C.POINTER(C.c_double)]),
        "bh_measure_fp32x2_rate": (C.c_int, [i32, C.POINTER(C.c_double)]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    lib._protos = protos
    _lib = lib
    return lib


def declared_symbols():
    """Function names declared in include/bhstep.h (used by the ABI test)."""
    import re
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bh_[a-z_0-9]+)\s*\(", text)))
