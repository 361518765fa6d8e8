"""Run the reference's own kernel sources under oracle/clshim -- TEST INFRASTRUCTURE ONLY.

Needs /root/reference (so: this container, not the GPU box).  Used by
tests/golden/make_golden.py to generate the committed golden vectors and by
tests/test_oracle_vs_reference.py when the reference tree is present.
"""
from __future__ import annotations

import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
NAMES = ["posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "step", "blockCount", "bodyCount",
         "radius", "maxDepth", "bottom", "mass", "child", "start", "sorted", "error"]
FLOATS = {"posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "radius", "mass"}
KERNELS = ["boundingBox", "buildTree", "summarizeTree", "sort", "calculateForce", "integrate"]


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "kernels", "nbody"))


def number_of_nodes(n: int) -> int:
    m = max(2 * n, 16384)
    return m + (-m) % 16


def build(n: int, fma: bool = False, theta05: bool = False) -> str:
    tag = ("fma" if fma else "nofma") + ("_theta05" if theta05 else "")
    exe = os.path.join(HERE, "_ref", "ref_step_n%d_%s" % (n, tag))
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "clshim"), "N=%d" % n, "FMA=%d" % int(fma),
                           "THETA05=%d" % int(theta05), "REF=" + REF], stdout=subprocess.DEVNULL)
    return exe


def run(arrays, steps: int = 1, stop_after: str = "integrate", fma: bool = False, theta05: bool = False):
    """Returns the reference's 20 buffers after kernel `stop_after` of step `steps`."""
    n = int(np.asarray(arrays[0]).size)
    exe = build(n, fma, theta05)
    m1 = number_of_nodes(n) + 1
    with tempfile.TemporaryDirectory() as tmp:
        inp, out = os.path.join(tmp, "u.bin"), os.path.join(tmp, "o.bin")
        with open(inp, "wb") as f:
            for a in arrays:
                f.write(np.ascontiguousarray(a, dtype=np.float32)[:n].tobytes())
        rc = subprocess.call([exe, inp, out, str(steps), str(KERNELS.index(stop_after))], stdout=subprocess.DEVNULL)
        if rc not in (0, 3):
            raise RuntimeError("reference run failed with exit code %d" % rc)
        raw = np.fromfile(out, dtype=np.uint32)
    bufs, off = {}, 0
    for name in NAMES:
        ln = 1 if name in ("step", "blockCount", "radius", "maxDepth", "bottom", "error") else (8 * m1 if name == "child" else m1)
        chunk = raw[off:off + ln]
        bufs[name] = chunk.view(np.float32 if name in FLOATS else np.int32).copy()
        off += ln
    assert off == raw.size
    return bufs
