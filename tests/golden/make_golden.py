#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE'S OWN
kernel sources (/root/reference/kernels/nbody/*.cl) under oracle/clshim.

Run in the build container (needs /root/reference); the .npz files are committed,
so the GPU box and CI never need the reference tree:

    python tests/golden/make_golden.py [--only NAME]

Every fixture holds the inputs (or names the bundled-universe input fixture),
the reference parameters, and the reference's buffers after calculateForce and
after integrate of the last step, with cells in canonical (DFS-from-root) order
because raw cell numbers are a race in the reference (buildtree.cl:109).
The reference is built without FMA contraction (-ffp-contract=off) and with
rsqrt(x) = 1/sqrt(x); THETA is the shipped (1.5f) unless `theta05` is set, in
which case the author's commented-out `THETA (0.5f * 0.5f)` line is used
(oracle/clshim/Makefile).
"""
import argparse
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import refshim  # noqa: E402
from gpu_nbody_b200 import universe as U  # noqa: E402


def bundled(name):
    d = np.load(os.path.join(HERE, name + ".npz"))
    n = d["x"].size
    z = np.zeros(n, dtype=np.float32)
    return [d["x"], d["y"], d["z"], z, z.copy(), z.copy(), np.full(n, d["mass"][0], dtype=np.float32)]


FIXTURES = {
    # name: (input, theta05, steps, store_inputs, full)
    "ref_twobody": (lambda: U.generate_arrays(U.TwoBodyUniverse(), 2), False, 1, True, True),
    "ref_eightbody": (lambda: U.generate_arrays(U.EightBodyUniverse(), 8), False, 1, True, True),
    "ref_bigtree": (lambda: U.generate_arrays(U.BigTreeUniverse(), 4), False, 1, True, True),
    "ref_plummer1024_theta05": (lambda: U.generate_arrays(U.PlummerUniverseGenerator(5), 1024), True, 1, True, True),
    "ref_plummer1000_ragged_theta05": (lambda: U.generate_arrays(U.PlummerUniverseGenerator(6), 1008), True, 2, True, True),
    "ref_disks2048_shipped_theta": (lambda: U.generate_arrays(U.TwoDiskGalaxiesGenerator(45, 46), 2048), False, 1, True, True),
    "ref_plummer4096_theta05_3steps": (lambda: U.generate_arrays(U.PlummerUniverseGenerator(5), 4096), True, 3, True, True),
    "ref_sphericaluniverse1_theta05": (lambda: bundled("sphericaluniverse1"), True, 1, False, False),
    "ref_montecarlouniverse1_shipped_theta": (lambda: bundled("montecarlouniverse1"), False, 1, False, False),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make(name):
    inp, theta05, steps, store_inputs, full = FIXTURES[name]
    arrays = inp()
    n = arrays[0].size
    m = refshim.number_of_nodes(n)
    force = refshim.run(arrays, steps=steps, stop_after="calculateForce", fma=False, theta05=theta05)
    integ = refshim.run(arrays, steps=steps, stop_after="integrate", fma=False, theta05=theta05)
    assert force["error"][0] == 0
    order, canon = oracle.canonicalize(force["child"], n, m)
    out = {
        "n": np.int32(n), "m": np.int32(m), "steps": np.int32(steps), "theta_macro": np.float32(0.25 if theta05 else 1.5),
        "eps2": np.float32(0.0025), "dt": np.float32(0.025),
        "root": np.array([force[k][m] for k in ("posX", "posY", "posZ")], dtype=np.float32),  # overwritten by the COM
        "radius": force["radius"], "bottom": force["bottom"], "maxDepth": force["maxDepth"], "step": force["step"],
        "cells": np.int32(order.size),
        "canon_child_sha256": np.array(sha(canon)), "sorted_sha256": np.array(sha(force["sorted"][:n])),
        "bodyCount_sha256": np.array(sha(force["bodyCount"][order])), "start_sha256": np.array(sha(force["start"][order])),
    }
    if store_inputs:
        for k, a in zip(("in_x", "in_y", "in_z", "in_vx", "in_vy", "in_vz", "in_mass"), arrays):
            out[k] = np.asarray(a, dtype=np.float32)
    sel_b = np.arange(n) if full else np.arange(0, n, 8)           # bodies kept
    sel_c = np.arange(order.size) if full else np.arange(0, order.size, 16)  # canonical cells kept
    out["body_sel"], out["cell_sel"] = sel_b.astype(np.int32), sel_c.astype(np.int32)
    if full:
        out["canon_child"], out["sorted"] = canon, force["sorted"][:n]
        out["bodyCount"], out["start"] = force["bodyCount"][order], force["start"][order]
    for k in ("posX", "posY", "posZ", "mass"):
        out["cell_" + k] = force[k][order][sel_c]
    for k in ("accX", "accY", "accZ", "velX", "velY", "velZ"):
        out["force_" + k] = force[k][:n][sel_b]
    for k in ("posX", "posY", "posZ", "velX", "velY", "velZ"):
        out["integ_" + k] = integ[k][:n][sel_b]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n", n, "cells", order.size, "maxDepth", int(force["maxDepth"][0]),
          "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")), flush=True)


def make_trajectory():
    """BASELINE configs[0]: 10 steps of the bundled universe with the reference kernels (about 4 minutes)."""
    arrays = bundled("sphericaluniverse1")
    n = arrays[0].size
    ref = refshim.run(arrays, steps=10, stop_after="integrate", fma=False, theta05=True)
    sel = np.arange(0, n, 8)
    np.savez_compressed(os.path.join(HERE, "ref_sphericaluniverse1_theta05_10steps.npz"), n=np.int32(n), steps=np.int32(10),
                        theta_macro=np.float32(0.25), eps2=np.float32(0.0025), dt=np.float32(0.025), body_sel=sel.astype(np.int32),
                        step=ref["step"], maxDepth=ref["maxDepth"], error=ref["error"],
                        **{"integ_" + k: ref[k][:n][sel] for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ")})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    if not refshim.available():
        raise SystemExit("needs /root/reference")
    for nm in FIXTURES:
        if a.only is None or a.only == nm:
            make(nm)
    if a.only in (None, "trajectory"):
        make_trajectory()
