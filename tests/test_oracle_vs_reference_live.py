"""CPU, only where /root/reference exists (the build container): run the reference's own kernel sources under
oracle/clshim *now* -- without and with compiler FMA contraction (the reference builds with MAD enabled) -- and hold
the oracle to them in the matching fma_policy.  Skipped on the GPU box, where the committed golden vectors stand in."""
import numpy as np
import pytest

import oracle
from oracle import refshim
from gpu_nbody_b200 import universe as U

pytestmark = pytest.mark.skipif(not refshim.available(), reason="needs the reference tree at /root/reference")


def _compare(arrays, fma, theta05, steps=1):
    n = arrays[0].size
    ref = refshim.run(arrays, steps=steps, stop_after="integrate", fma=fma, theta05=theta05)
    o = oracle.OracleSim(n, *arrays, theta_macro=(0.25 if theta05 else 1.5), fma_policy=int(fma))
    assert o.step(steps) == 0 and ref["error"][0] == 0
    ro, rc = oracle.canonicalize(ref["child"], n, o.m)
    oo, oc = oracle.canonicalize(o.child, n, o.m)
    # integer work: bit-exact
    assert np.array_equal(rc, oc) and np.array_equal(ref["sorted"][:n], o.sorted[:n])
    assert np.array_equal(ref["bodyCount"][ro], o.bodyCount[oo]) and np.array_equal(ref["start"][ro], o.start[oo])
    assert ref["bottom"][0] == o.bottom[0] and ref["maxDepth"][0] == o.maxDepth[0] and ref["step"][0] == o.buf["step"][0]
    assert ref["radius"].view(np.uint32)[0] == o.radius.view(np.uint32)[0]
    assert np.array_equal(ref["mass"][ro].view(np.uint32), o.mass[oo].view(np.uint32)) or np.allclose(ref["mass"][ro], o.mass[oo], rtol=3e-7)
    # float work: the reference's own summation order in summarise is timing dependent -> ulp-level differences
    a = np.stack([ref[k][:n] for k in ("accX", "accY", "accZ")], 1).astype(np.float64)
    b = np.stack([o.buf[k][:n] for k in ("accX", "accY", "accZ")], 1).astype(np.float64)
    err = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-30)
    assert err.max() < 1e-5
    for k in ("posX", "posY", "posZ", "velX", "velY", "velZ"):
        np.testing.assert_allclose(ref[k][:n], o.buf[k][:n], rtol=1e-6, atol=1e-7)
    return ref, o


@pytest.mark.parametrize("fma", [False, True])
def test_single_cell_universe_is_bit_exact(fma):
    arrays = U.generate_arrays(U.EightBodyUniverse(), 8)
    ref, o = _compare(arrays, fma, theta05=False)
    for k in ("accX", "accY", "accZ", "posX", "velX"):
        assert np.array_equal(ref[k][:8].view(np.uint32), o.buf[k][:8].view(np.uint32)), k


@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("theta05", [False, True])
def test_plummer_1024(fma, theta05):
    _compare(U.generate_arrays(U.PlummerUniverseGenerator(5), 1024), fma, theta05)


def test_uniform_2048_two_steps_with_contraction():
    _compare(U.generate_arrays(U.RandomCubicUniverseGenerator(6.0, 9), 2048), True, True, steps=2)
