"""Compare a buffer set (CPU oracle or CUDA path) with the golden vectors that
tests/golden/make_golden.py produced from the reference's own kernel sources."""
import hashlib
import os

import numpy as np

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REFERENCE_FIXTURES = ["ref_twobody", "ref_eightbody", "ref_bigtree", "ref_plummer1024_theta05", "ref_plummer1000_ragged_theta05",
                      "ref_disks2048_shipped_theta", "ref_plummer4096_theta05_3steps", "ref_sphericaluniverse1_theta05",
                      "ref_montecarlouniverse1_shipped_theta"]
BUNDLED = {"ref_sphericaluniverse1_theta05": "sphericaluniverse1", "ref_montecarlouniverse1_shipped_theta": "montecarlouniverse1"}

ACC_RTOL = 1e-4      # north_star: per-body acceleration within 1e-4 relative
COM_ATOL_RADII = 2e-6  # centre of mass: the reference's own summation order is timing dependent (summarizetree.cl:93-150)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def bundled_inputs(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = d["x"].size
    z = np.zeros(n, dtype=np.float32)
    return [d["x"], d["y"], d["z"], z, z.copy(), z.copy(), np.full(n, d["mass"][0], dtype=np.float32)]


def inputs(name, g):
    if name in BUNDLED:
        return bundled_inputs(BUNDLED[name])
    return [g[k] for k in ("in_x", "in_y", "in_z", "in_vx", "in_vy", "in_vz", "in_mass")]


def check_after_force(buf, g):
    """buf: name -> full logical buffer (reference conventions) after calculateForce of the last step."""
    n, m = int(g["n"]), int(g["m"])
    assert int(np.asarray(buf["bottom"]).ravel()[0]) == int(g["bottom"][0])
    assert int(np.asarray(buf["maxDepth"]).ravel()[0]) == int(g["maxDepth"][0])
    assert int(np.asarray(buf["step"]).ravel()[0]) == int(g["step"][0])
    assert np.float32(np.asarray(buf["radius"]).ravel()[0]).view(np.uint32) == g["radius"].view(np.uint32)[0], "radius bits"
    order, canon = oracle.canonicalize(buf["child"], n, m)
    assert order.size == int(g["cells"])
    # integer work: bit-exact against the reference
    assert sha(canon) == str(g["canon_child_sha256"]), "canonical child structure differs from the reference"
    assert sha(np.asarray(buf["sorted"][:n], dtype=np.int32)) == str(g["sorted_sha256"]), "sorted[] differs from the reference"
    assert sha(np.asarray(buf["bodyCount"], dtype=np.int32)[order]) == str(g["bodyCount_sha256"]), "bodyCount differs"
    assert sha(np.asarray(buf["start"], dtype=np.int32)[order]) == str(g["start_sha256"]), "start differs"
    if "canon_child" in g:
        assert np.array_equal(canon, g["canon_child"]) and np.array_equal(buf["sorted"][:n], g["sorted"])
    # float work
    radius = float(g["radius"][0])
    csel, bsel = g["cell_sel"], g["body_sel"]
    assert np.array_equal(np.asarray(buf["mass"])[order][csel].view(np.uint32) != 0, g["cell_mass"].view(np.uint32) != 0)
    np.testing.assert_allclose(np.asarray(buf["mass"])[order][csel], g["cell_mass"], rtol=2e-6, atol=0)
    for k in ("posX", "posY", "posZ"):
        np.testing.assert_allclose(np.asarray(buf[k])[order][csel], g["cell_" + k], rtol=0, atol=COM_ATOL_RADII * radius,
                                   err_msg="centre of mass " + k)
    a = np.stack([np.asarray(buf[k])[:n][bsel] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    r = np.stack([g["force_" + k] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    err = np.linalg.norm(a - r, axis=1) / np.maximum(np.linalg.norm(r, axis=1), 1e-30)
    assert err.max() <= ACC_RTOL, "acceleration vs reference: max rel err %g" % err.max()
    for k in ("velX", "velY", "velZ"):
        np.testing.assert_allclose(np.asarray(buf[k])[:n][bsel], g["force_" + k], rtol=1e-5, atol=1e-6)
    return float(err.max())


def check_after_integrate(buf, g):
    n = int(g["n"])
    bsel = g["body_sel"]
    for k in ("posX", "posY", "posZ", "velX", "velY", "velZ"):
        np.testing.assert_allclose(np.asarray(buf[k])[:n][bsel], g["integ_" + k], rtol=1e-5, atol=1e-6, err_msg="integrate " + k)


TRAJECTORY_FIXTURE = "ref_sphericaluniverse1_theta05_10steps"  # BASELINE configs[0]: the bundled universe, theta = 0.5, 10 steps


def check_trajectory(buf, g):
    """State after `steps` full steps against the reference kernels' state (every 8th body)."""
    n = int(g["n"])
    sel = g["body_sel"]
    assert int(np.asarray(buf["step"]).ravel()[0]) == int(g["step"][0]) and int(np.asarray(buf["maxDepth"]).ravel()[0]) == int(g["maxDepth"][0])
    for k in ("posX", "posY", "posZ", "velX", "velY", "velZ"):
        np.testing.assert_allclose(np.asarray(buf[k])[:n][sel], g["integ_" + k], rtol=0, atol=1e-6, err_msg=k)
    a = np.stack([np.asarray(buf[k])[:n][sel] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    r = np.stack([g["integ_" + k] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    err = np.linalg.norm(a - r, axis=1) / np.maximum(np.linalg.norm(r, axis=1), 1e-30)
    assert err.max() <= ACC_RTOL, "acceleration after %d steps vs reference: max rel err %g" % (int(g["steps"]), err.max())
    return float(err.max())
