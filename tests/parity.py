"""Shared helpers for the parity tests: run the CUDA path (through the C ABI) and
the CPU oracle on the same inputs and compare stage by stage.

Parity classes:
  * integer work -- canonical child structure, bodyCount, start, sorted, bottom,
    maxDepth, interaction / open counts: bit-exact;
  * fp32 produced by exactly rounded operations -- root box, cell centres,
    centres of mass, integrate, velocity correction: bit-exact under the shared
    FMA policy;
  * accelerations (rsqrt, summation order): relative error <= 1e-4 per body.
"""
from __future__ import annotations

import numpy as np

import oracle
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode
from gpu_nbody_b200.universe import ArrayUniverseGenerator

ACC_RTOL = 1e-4  # BASELINE.json north_star: per-body acceleration within 1e-4 relative


def make_pair(arrays, theta=0.5, eps2=0.0025, dt=0.025, vote_width=16, theta_macro=None, counting=True):
    n = int(np.asarray(arrays[0]).size)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, ArrayUniverseGenerator(*arrays), theta=theta, eps2=eps2, dt=dt,
                                      vote_width=vote_width, theta_macro=theta_macro)
    sim.init(None)
    if counting:
        sim.setCounting(True)
    orc = oracle.OracleSim(n, *arrays, theta=theta, eps2=eps2, dt=dt, vote_width=vote_width, fma_policy=1,
                           theta_macro=theta_macro)
    return sim, orc


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(a, b, what):
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    bad = np.nonzero(bits(a) != bits(b))[0]
    assert bad.size == 0, "%s: %d of %d values differ bitwise, first at %d: %r vs %r" % (
        what, bad.size, a.size, bad[0], a[bad[0]], b[bad[0]])


def sync_oracle_from_gpu(sim, orc):
    """Make the oracle start the next step from the CUDA path's state (positions,
    velocities, accelerations, step, maxDepth), so that per-stage comparisons
    stay exact after accelerations have diverged by rounding."""
    n = sim.nbodies
    for name in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "mass"):
        orc.buf[name][:n] = sim.readBuffer(name, n)
    orc.buf["step"][0] = sim.scalar("step")
    orc.buf["maxDepth"][0] = sim.scalar("maxDepth")


def check_bounding_box(sim, orc):
    sim.boundingBox(); orc.bounding_box()
    m = orc.m
    for name in ("posX", "posY", "posZ"):
        assert_bits_equal(sim.readBuffer(name)[m], orc.buf[name][m], "root " + name)
    assert_bits_equal(sim.scalar("radius"), orc.radius[0], "radius")
    assert sim.scalar("bottom") == m == orc.bottom[0]
    assert sim.scalar("step") == orc.buf["step"][0]
    assert sim.scalar("blockCount") == 0
    assert sim.readBuffer("mass")[m] == -1.0 and sim.readBuffer("start")[m] == 0
    assert (sim.readBuffer("child")[8 * m:] == -1).all()


def _canon_both(sim, orc):
    n, m = orc.n, orc.m
    gchild = sim.readBuffer("child")
    gorder, gcanon = oracle.canonicalize(gchild, n, m)
    oorder, ocanon = oracle.canonicalize(orc.child, n, m)
    assert gorder.size == oorder.size, "cell count %d vs %d" % (gorder.size, oorder.size)
    assert np.array_equal(gcanon, ocanon), "canonical child structure differs"
    return gorder, oorder


def check_build_tree(sim, orc):
    sim.buildTree(); rc = orc.build_tree()
    assert rc == 0
    assert sim.scalar("bottom") == orc.bottom[0]
    assert sim.scalar("maxDepth") == orc.maxDepth[0]
    gorder, oorder = _canon_both(sim, orc)
    assert gorder.size == orc.cells_used
    for name in ("posX", "posY", "posZ"):  # geometric cell centres, buildtree.cl:134-136
        assert_bits_equal(sim.readBuffer(name)[gorder], orc.buf[name][oorder], "cell centre " + name)
    return gorder, oorder


def check_summarize(sim, orc):
    sim.summarizeTree(); orc.summarize()
    gorder, oorder = _canon_both(sim, orc)  # now compacted (summarizetree.cl:77-81)
    assert np.array_equal(sim.readBuffer("bodyCount")[gorder], orc.bodyCount[oorder]), "bodyCount differs"
    for name in ("posX", "posY", "posZ", "mass"):
        assert_bits_equal(sim.readBuffer(name)[gorder], orc.buf[name][oorder], "cell " + name)
    assert sim.readBuffer("bodyCount")[orc.m] == orc.n
    return gorder, oorder


def check_sort(sim, orc, gorder, oorder):
    sim.sort(); orc.sort()
    n = orc.n
    gs = sim.readBuffer("sorted", n)
    assert np.array_equal(gs, orc.sorted[:n]), "sorted[] differs"
    assert np.array_equal(np.sort(gs), np.arange(n)), "sorted[] is not a permutation"
    assert np.array_equal(sim.readBuffer("start")[gorder], orc.start[oorder]), "start differs"


def rel_acc_error(sim, orc):
    n = orc.n
    ga = np.stack([sim.readBuffer(k, n) for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    oa = np.stack([orc.buf[k][:n] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    num = np.linalg.norm(ga - oa, axis=1)
    den = np.linalg.norm(oa, axis=1)
    return num / np.maximum(den, 1e-30)


def check_force(sim, orc):
    n = orc.n
    step = int(sim.scalar("step"))
    vel0 = [sim.readBuffer(k, n) for k in ("velX", "velY", "velZ")]
    acc0 = [sim.readBuffer(k, n) for k in ("accX", "accY", "accZ")]
    sim.calculateForce(); rc = orc.calculate_force()
    assert rc == 0
    st = sim.stats()
    # the *set* of (vote group, node) interactions is integer work: exact
    assert st["interactions"] == orc.interactions, "interactions %d vs %d" % (st["interactions"], orc.interactions)
    assert st["opens"] == orc.opens, "opens %d vs %d" % (st["opens"], orc.opens)
    err = rel_acc_error(sim, orc)
    assert err.max() <= ACC_RTOL, "max relative acceleration error %g > %g" % (err.max(), ACC_RTOL)
    # velocity correction (calculateforce.cl:174-179) from the CUDA path's own accelerations: exact
    dt = np.float32(orc.state.timestep)
    for k, v0, a0 in zip("XYZ", vel0, acc0):
        a1 = sim.readBuffer("acc" + k, n)
        v1 = sim.readBuffer("vel" + k, n)
        want = v0 + ((a1 - a0) * dt) * np.float32(0.5) if step > 0 else v0
        assert_bits_equal(v1, want, "velocity correction vel" + k)
    return err


def check_integrate(sim, orc):
    """integrate.cl:29-43 is exactly rounded work: run the oracle's integrate on
    the CUDA path's own pre-integrate state and compare bits."""
    n = orc.n
    for name in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ"):
        orc.buf[name][:n] = sim.readBuffer(name, n)
    sim.integrate(); orc.integrate()
    for name in ("posX", "posY", "posZ", "velX", "velY", "velZ"):
        assert_bits_equal(sim.readBuffer(name, n), orc.buf[name][:n], "integrate " + name)


def check_full_step(sim, orc):
    check_bounding_box(sim, orc)
    check_build_tree(sim, orc)
    gorder, oorder = check_summarize(sim, orc)
    check_sort(sim, orc, gorder, oorder)
    err = check_force(sim, orc)
    check_integrate(sim, orc)
    return err
