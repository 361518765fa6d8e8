"""CPU: the oracle (oracle/bh_oracle.c) against golden vectors produced by the
reference's own kernel sources (tests/golden/make_golden.py), in both FMA policies,
plus the survey's known answers for the hand-written universes."""
import numpy as np
import pytest

import goldencheck as gc
import oracle
from gpu_nbody_b200 import universe as U


def run_oracle(arrays, g, fma):
    n = int(g["n"])
    o = oracle.OracleSim(n, *arrays, theta_macro=float(g["theta_macro"]), eps2=float(g["eps2"]), dt=float(g["dt"]), fma_policy=fma)
    for _ in range(int(g["steps"]) - 1):
        assert o.step(1) == 0
    o.bounding_box(); assert o.build_tree() == 0; o.summarize(); o.sort(); assert o.calculate_force() == 0
    return o


@pytest.mark.parametrize("fma", [0, 1])
@pytest.mark.parametrize("name", gc.REFERENCE_FIXTURES)
def test_oracle_matches_reference_kernels(name, fma):
    g = gc.load(name)
    o = run_oracle(gc.inputs(name, g), g, fma)
    gc.check_after_force(o.buf, g)
    o.integrate()
    gc.check_after_integrate(o.buf, g)


def test_oracle_bitwise_where_the_reference_is_deterministic():
    """Without contraction, with 1/sqrt for rsqrt and a single-child-order tree, oracle == reference to the bit."""
    for name in ("ref_twobody", "ref_eightbody", "ref_bigtree"):
        g = gc.load(name)
        o = run_oracle(gc.inputs(name, g), g, 0)
        n = int(g["n"])
        for k in ("accX", "accY", "accZ"):
            assert np.array_equal(o.buf[k][:n].view(np.uint32), g["force_" + k].view(np.uint32)), (name, k)
        order, _ = oracle.canonicalize(o.child, n, o.m)
        for k in ("posX", "posY", "posZ", "mass"):
            assert np.array_equal(o.buf[k][order].view(np.uint32), g["cell_" + k].view(np.uint32)), (name, k)


def test_survey_known_answers():
    """SURVEY.md appendix D."""
    o = oracle.OracleSim(2, *U.generate_arrays(U.TwoBodyUniverse(), 2))
    o.bounding_box(); o.build_tree()
    m = o.m
    assert (o.posX[m], o.posY[m], o.posZ[m]) == (np.float32(-0.050000012), np.float32(0.050000012), 0.0)
    assert o.radius[0] == np.float32(1.05) and m == 16384
    assert list(o.child[8 * m:8 * m + 8]) == [0, -1, -1, -1, -1, -1, -1, 1]
    o.summarize(); o.sort()
    assert list(o.child[8 * m:8 * m + 8]) == [0, 1, -1, -1, -1, -1, -1, -1] and list(o.sorted[:2]) == [0, 1]
    assert o.maxDepth[0] == 1 and o.mass[m] == 0.25

    o = oracle.OracleSim(8, *U.generate_arrays(U.EightBodyUniverse(), 8))
    o.bounding_box(); o.build_tree(); o.summarize(); o.sort()
    m = o.m
    assert list(o.child[8 * m:8 * m + 8]) == [0, 1, 2, 4, 3, 6, 5, 7] == list(o.sorted[:8])
    assert o.radius[0] == np.float32(1.1) and o.mass[m] == 1.0
    assert (o.posX[m], o.posY[m], o.posZ[m]) == (np.float32(-0.012500003), np.float32(0.025000006), 0.0)

    o = oracle.OracleSim(4, *U.generate_arrays(U.BigTreeUniverse(), 4))
    o.bounding_box(); o.build_tree()
    m = o.m
    assert list(o.child[8 * m:8 * m + 8]) == [m - 1, 1, -1, -1, -1, -1, -1, 2]
    assert list(o.child[8 * (m - 1):8 * m]) == [3, -1, -1, -1, -1, -1, -1, 0]
    o.summarize(); o.sort()
    assert list(o.sorted[:4]) == [3, 0, 1, 2] and o.maxDepth[0] == 2 and o.cells_used == 2
    assert abs(o.mass[m] - 0.3) < 1e-7 and abs(o.posX[m] - 3333.3333) < 1e-3


def test_bundled_universe_statistics():
    """SURVEY.md appendix D, sphericaluniverse1 at theta = 0.5."""
    a = gc.bundled_inputs("sphericaluniverse1")
    o = oracle.OracleSim(32768, *a)
    assert o.step(1) == 0
    assert o.cells_used == 23562 and o.bottom[0] == 41975 and o.maxDepth[0] == 13
    assert abs(o.interactions / 32768 - 375) < 1 and abs(o.opens / 32768 - 140) < 1
    assert o.radius[0] == np.float32(0.9999582767486572)


def test_tree_is_insertion_order_independent():
    a = U.generate_arrays(U.PlummerUniverseGenerator(9), 5000)
    ref = oracle.OracleSim(5000, *a); ref.bounding_box(); ref.build_tree(); ref.summarize(); ref.sort()
    _, canon_ref = ref.canonical()
    perm = np.random.default_rng(0).permutation(5000)
    o = oracle.OracleSim(5000, *[x[perm] for x in a]); o.bounding_box(); o.build_tree(); o.summarize(); o.sort()
    # map permuted body ids back to the original numbering
    _, canon = o.canonical()
    body = (canon >= 0) & (canon < 5000)
    canon[body] = perm[canon[body]]
    assert np.array_equal(canon, canon_ref)
    assert np.array_equal(perm[o.sorted[:5000]], ref.sorted[:5000])
    assert o.maxDepth[0] == ref.maxDepth[0] and o.bottom[0] == ref.bottom[0]


def test_overflow_and_depth_errors():
    a = U.generate_arrays(U.PlummerUniverseGenerator(1), 64)
    for k in range(3):
        a[k][10] = a[k][3]  # coincident bodies: buildtree.cl:112-119
    o = oracle.OracleSim(64, *a)
    assert o.step(1) == 1 and o.error[0] == 1 and o.bottom[0] == o.m


@pytest.mark.parametrize("fma", [0, 1])
def test_oracle_ten_steps_of_the_bundled_universe(fma):
    """BASELINE configs[0]: sphericaluniverse1, theta = 0.5, 10 steps, against the reference kernels' own 10 steps."""
    g = gc.load(gc.TRAJECTORY_FIXTURE)
    o = oracle.OracleSim(32768, *gc.bundled_inputs("sphericaluniverse1"), theta_macro=float(g["theta_macro"]), eps2=float(g["eps2"]),
                         dt=float(g["dt"]), fma_policy=fma)
    assert o.step(10) == 0
    gc.check_trajectory(o.buf, g)


@pytest.mark.parametrize("n", [1, 2, 17, 1000, 50000])
def test_parallel_build_gives_the_same_tree(n):
    """bho_build_tree_parallel (concurrent insertion with CAS-locked child slots, buildtree.cl:93-180 as the reference runs
    it; the CPU baseline of bench.py) builds the tree of the sequential restatement: same canonical structure, same
    cell count and depth, hence the same sorted order and bit-identical centres of mass."""
    import oracle
    from gpu_nbody_b200 import universe as U
    a = U.generate_arrays(U.TwoDiskGalaxiesGenerator(7, 8) if n > 100 else U.PlummerUniverseGenerator(7), n)
    oracle.lib().bho_set_num_threads(4)
    seq, par = oracle.OracleSim(n, *a), oracle.OracleSim(n, *a)
    for o, build in ((seq, seq.build_tree), (par, par.build_tree_parallel)):
        o.bounding_box()
        assert build() == 0
    assert seq.bottom[0] == par.bottom[0] and seq.maxDepth[0] == par.maxDepth[0]
    so, sc = oracle.canonicalize(seq.child, n, seq.m)
    po, pc = oracle.canonicalize(par.child, n, par.m)
    assert np.array_equal(sc, pc)
    for o in (seq, par):
        o.summarize(); o.sort()
    assert np.array_equal(seq.sorted[:n], par.sorted[:n])
    for k in ("posX", "posY", "posZ", "mass"):
        assert np.array_equal(seq.buf[k][so].view(np.uint32), par.buf[k][po].view(np.uint32))
    # coincident bodies exhaust the pool in both (buildtree.cl:112-119)
    if n >= 17:
        for k in range(3):
            a[k][5] = a[k][3]
        bad = oracle.OracleSim(n, *a)
        bad.bounding_box()
        assert bad.build_tree_parallel() == 1 and bad.error[0] == 1
