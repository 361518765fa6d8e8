"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs (run with ``pytest -m gpu`` on the B200 box)."""
import os

import numpy as np
import pytest

import parity
from gpu_nbody_b200 import BhError, GPUBarnesHutNBodySimulation, Mode, universe as U

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def bundled(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = d["x"].size
    z = np.zeros(n, dtype=np.float32)
    return [d["x"], d["y"], d["z"], z, z.copy(), z.copy(), np.full(n, d["mass"][0], dtype=np.float32)]


def gen(generator, n):
    return U.generate_arrays(generator, n)


# ---- the reference's hand-written test universes (universe/test/*.java) ----------------
@pytest.mark.parametrize("generator,n", [(U.TwoBodyUniverse(), 2), (U.EightBodyUniverse(), 8), (U.BigTreeUniverse(), 4)])
def test_reference_test_universes(generator, n):
    sim, orc = parity.make_pair(gen(generator, n))
    parity.check_full_step(sim, orc)
    sim.close()


@pytest.mark.parametrize("n", [1, 3, 16, 17, 100, 1000, 4096])
def test_small_and_ragged_sizes(n):
    """N below one vote group, not a multiple of 16/32, around one CTA."""
    sim, orc = parity.make_pair(gen(U.PlummerUniverseGenerator(seed=n), n))
    parity.check_full_step(sim, orc)
    sim.close()


@pytest.mark.parametrize("name", ["sphericaluniverse1", "montecarlouniverse1"])
def test_bundled_universe_three_steps(name):
    """BASELINE config 1: the bundled 32768-body universes, theta = 0.5."""
    sim, orc = parity.make_pair(bundled(name))
    for _ in range(3):
        parity.sync_oracle_from_gpu(sim, orc)
        parity.check_full_step(sim, orc)
    assert sim.scalar("step") == 2
    sim.close()


def test_bundled_universe_shipped_theta():
    """The THETA (1.5f) the reference ships with (calculateforce.cl:16)."""
    sim, orc = parity.make_pair(bundled("sphericaluniverse1"), theta_macro=1.5)
    parity.check_full_step(sim, orc)
    sim.close()


@pytest.mark.parametrize("generator,n", [
    (U.PlummerUniverseGenerator(42), 262144),
    (U.RandomCubicUniverseGenerator(6.0, 44), 131072),
    (U.TwoDiskGalaxiesGenerator(45, 46), 131072),
    (U.SphericalUniverseGenerator(47), 50000),
])
def test_distributions(generator, n):
    sim, orc = parity.make_pair(gen(generator, n))
    for _ in range(2):
        parity.sync_oracle_from_gpu(sim, orc)
        parity.check_full_step(sim, orc)
    sim.close()


@pytest.mark.parametrize("theta", [0.3, 0.8])
def test_theta_sweep(theta):
    sim, orc = parity.make_pair(gen(U.TwoDiskGalaxiesGenerator(45, 46), 65536), theta=theta)
    parity.check_full_step(sim, orc)
    sim.close()


def test_vote_width_32_mode():
    """The non-reference 32-wide vote: still the oracle's semantics at vote_width = 32."""
    sim, orc = parity.make_pair(gen(U.PlummerUniverseGenerator(7), 32768), vote_width=32)
    parity.check_full_step(sim, orc)
    sim.close()


def test_insertion_order_does_not_change_the_tree():
    arrays = gen(U.PlummerUniverseGenerator(11), 65536)
    results = []
    for mode in (0, 1):
        sim, orc = parity.make_pair(arrays)
        sim.setInsertionOrder(mode)
        sim.step(2)           # second step inserts in the first step's sorted order when mode = 1
        parity.sync_oracle_from_gpu(sim, orc)
        parity.check_full_step(sim, orc)
        results.append([sim.readBuffer(k, 65536) for k in ("posX", "velX", "sorted")])
        sim.close()
    for a, b in zip(*results):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_step_equals_stage_sequence_and_is_deterministic():
    """bh_step(k) == k x the six stage calls (GPUBH:258-263); run to run bit-identical."""
    arrays = gen(U.PlummerUniverseGenerator(5), 20000)
    outs = []
    for use_step in (True, False, True):
        sim, _ = parity.make_pair(arrays, counting=False)
        if use_step:
            sim.step(3)
        else:
            for _ in range(3):
                sim.boundingBox(); sim.buildTree(); sim.summarizeTree(); sim.sort(); sim.calculateForce(); sim.integrate()
        outs.append([sim.readBuffer(k, 20000) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "sorted")])
        assert sim.scalar("step") == 2
        sim.close()
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_energy_drift_100_steps_matches_oracle():
    """north_star: matched total-energy drift over 100 steps at the same theta and dt."""
    import oracle
    n = 4096
    arrays = gen(U.PlummerUniverseGenerator(3), n)
    sim, orc = parity.make_pair(arrays, counting=False)
    e0 = sum(orc.energy())
    sim.step(100)
    assert orc.step(100) == 0
    g = [sim.readBuffer(k, n) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "mass")]
    eg = sum(oracle.energy(*g))
    eo = sum(orc.energy())
    drift_g, drift_o = (eg - e0) / abs(e0), (eo - e0) / abs(e0)
    # both integrate the same Hamiltonian with the same scheme; trajectories diverge
    # chaotically at the 1e-7 level, the energy error does not
    assert abs(drift_o) < 5e-3 and abs(drift_g) < 5e-3, (drift_g, drift_o)
    assert abs(drift_g - drift_o) < 5e-4, (drift_g, drift_o)
    sim.close()


def test_coincident_bodies_report_error_1():
    """buildtree.cl:112-119: identical positions exhaust the pool -> error = 1, no hang."""
    n = 64
    a = gen(U.PlummerUniverseGenerator(1), n)
    for k in range(3):
        a[k][10] = a[k][3]
    sim, _ = parity.make_pair(a, counting=False)
    with pytest.raises(BhError) as ei:
        sim.step(1)
    assert ei.value.code == 1
    assert sim.scalar("error") == 1
    sim.close()


def test_copy_vertices_and_read_conventions():
    n = 1000
    a = gen(U.PlummerUniverseGenerator(2), n)
    sim = GPUBarnesHutNBodySimulation(Mode.GL_INTEROP, n, U.ArrayUniverseGenerator(*a))
    sim.init(None)
    sim.initGLBuffers(None, -1, -1)
    assert sim.getNumberOfBodies() == n
    assert sim.scalar("step") == -1 and sim.scalar("maxDepth") == 1  # GPUBH:165,170
    sim.step()
    pos4, vel4 = sim.copyVertices()
    assert np.array_equal(pos4[:, 0], sim.readBuffer("posX", n)) and np.all(pos4[:, 3] == 1.0)
    assert np.array_equal(vel4[:, 2], sim.readBuffer("velZ", n)) and np.all(vel4[:, 3] == 1.0)
    m = sim.numberOfNodes
    assert sim.readBuffer("child").size == 8 * (m + 1) and sim.readBuffer("velX").size == m + 1
    assert not sim.readBuffer("velX")[n:].any() and not sim.readBuffer("sorted")[n:].any()
    sim.close()


def test_vertex_buffers_on_the_device():
    """copyvertices.cl:8-17 into DEVICE buffers (what a CUDA-mapped GL vertex buffer is): as a call
    (bh_copy_vertices_device) and fused into every step's finish pass (bh_set_vertex_buffers), checked against
    the oracle's positions -- not against bh_read."""
    import torch
    n = 5000
    a = gen(U.PlummerUniverseGenerator(13), n)
    sim, orc = parity.make_pair(a, counting=False)
    pos = torch.zeros((n, 4), device="cuda"); vel = torch.zeros((n, 4), device="cuda")
    torch.cuda.synchronize()
    sim.setVertexBuffers(pos.data_ptr(), vel.data_ptr())
    sim.step(2)
    assert orc.step(2) == 0
    want_p = np.stack([orc.buf[k][:n] for k in ("posX", "posY", "posZ")], 1)
    want_v = np.stack([orc.buf[k][:n] for k in ("velX", "velY", "velZ")], 1)
    p, v = pos.cpu().numpy(), vel.cpu().numpy()
    assert np.all(p[:, 3] == 1.0) and np.all(v[:, 3] == 1.0)
    assert np.allclose(p[:, :3], want_p, rtol=0, atol=1e-6) and np.allclose(v[:, :3], want_v, rtol=1e-5, atol=1e-6)
    sim.setVertexBuffers(0, 0)
    sim.step(1)
    assert np.array_equal(pos.cpu().numpy(), p)           # switched off: untouched
    assert orc.step(1) == 0
    pos2 = torch.zeros((n, 4), device="cuda")
    sim.copyVerticesDevice(pos2.data_ptr(), 0)
    sim._check(sim._lib.bh_check(sim.handle))
    want_p = np.stack([orc.buf[k][:n] for k in ("posX", "posY", "posZ")], 1)
    assert np.allclose(pos2.cpu().numpy()[:, :3], want_p, rtol=0, atol=2e-6)
    hp, hv = sim.copyVertices()                             # host destinations: same numbers
    assert np.array_equal(hp, pos2.cpu().numpy())
    sim.close()


def test_async_vertex_readback_overlapping_the_next_upload():
    """bh_copy_vertices_async + bh_wait_copies: the read-back of one step's vertices is still in flight while the next
    upload and step are enqueued (bench.py's e2e loop); what arrives is that step's state, not the next one's."""
    import torch
    n = 300_000
    a = gen(U.PlummerUniverseGenerator(31), n)
    b = gen(U.PlummerUniverseGenerator(32), n)
    sim, _ = parity.make_pair(a, counting=False)
    lib = sim._lib
    sim.step(1)
    want_p, want_v = [x.copy() for x in sim.copyVertices()]
    sim.upload(*a)
    sim.step(1)
    pos4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    vel4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    sim._check(lib.bh_copy_vertices_async(sim.handle, pos4.data_ptr(), vel4.data_ptr()))
    sim.upload(*b)          # a different universe goes up while the vertices come down
    sim.step(1)
    sim._check(lib.bh_wait_copies(sim.handle))
    assert np.array_equal(pos4.numpy(), want_p) and np.array_equal(vel4.numpy(), want_v)
    sim.close()


def test_async_upload_equals_upload():
    """bh_upload_async (velocities arrive on a second stream while the step's tree stages and walk already run) leaves
    the same state as bh_upload, whether a step or a read follows it."""
    import torch
    n = 200_000
    a = gen(U.TwoDiskGalaxiesGenerator(5, 6), n)
    ref, _ = parity.make_pair(a, counting=False)
    ref.step(2)
    want = {k: ref.readBuffer(k, n) for k in ("posX", "velX", "velZ", "accY", "sorted")}
    ref.close()
    pinned = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in a]
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
    sim.init(None)
    lib = sim._lib
    for _ in range(2):   # the second round overwrites a stepped state
        sim._check(lib.bh_upload_async(sim.handle, *(p.data_ptr() for p in pinned)))
        assert np.array_equal(sim.readBuffer("velY", n).view(np.uint32), a[4].view(np.uint32))   # a read waits for the velocities
        sim._check(lib.bh_upload_async(sim.handle, *(p.data_ptr() for p in pinned)))
        sim.step(2)
        for k, w in want.items():
            assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), w.view(np.uint32)), k
    sim.close()


def test_pipelined_host_loop():
    """bench.py's e2e loop: upload_async / step_async / copy_vertices_async with no wait between iterations (the next
    inputs cross the link while the current step computes); every iteration's vertices are those of a plain
    upload + step + copy_vertices of that iteration's input."""
    import torch
    n = 150_000
    universes = [gen(U.PlummerUniverseGenerator(40 + i), n) for i in range(3)]
    want = []
    ref = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
    ref.init(None)
    for a in universes:
        ref.upload(*a); ref.step(1)
        p, v = ref.copyVertices()
        want.append((p.copy(), v.copy()))
    ref.close()
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
    sim.init(None)
    lib = sim._lib
    pinned = [[torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in a] for a in universes]
    outs = [(torch.empty((n, 4), dtype=torch.float32).pin_memory(), torch.empty((n, 4), dtype=torch.float32).pin_memory()) for _ in universes]
    for rounds in range(2):
        for a, (p4, v4) in zip(pinned, outs):
            sim._check(lib.bh_upload_async(sim.handle, *(t.data_ptr() for t in a)))
            sim._check(lib.bh_step_async(sim.handle, 1))
            sim._check(lib.bh_copy_vertices_async(sim.handle, p4.data_ptr(), v4.data_ptr()))
        sim._check(lib.bh_wait_copies(sim.handle))
        sim._check(lib.bh_check(sim.handle))
        for (p4, v4), (wp, wv) in zip(outs, want):
            assert np.array_equal(p4.numpy(), wp) and np.array_equal(v4.numpy(), wv)
    sim.close()


def test_native_universe_writer(tmp_path):
    """bh_write_universe_file == UniverseSerializer.serialize of the current state: readable by the Python reader
    (same wire format as the reference's files, test_host.py) and by the native loader; a restart from the dump
    continues bit-identically."""
    n = 3000
    a = gen(U.PlummerUniverseGenerator(4), n)
    sim, _ = parity.make_pair(a, counting=False)
    sim.step(3)
    path = tmp_path / "dump.universe"
    sim.writeUniverseFile(path)
    assert path.stat().st_size == 93 + 28 * n
    cnt, back = U.read_universe(path)
    assert cnt == n
    for k, arr in zip(("posX", "posY", "posZ", "velX", "velY", "velZ", "mass"), back):
        assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), arr.view(np.uint32)), k
    other = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
    other.init(None)
    other.uploadUniverseFile(path)
    for k, arr in zip(("posX", "velY", "mass"), (back[0], back[4], back[6])):
        assert np.array_equal(other.readBuffer(k, n).view(np.uint32), arr.view(np.uint32))
    other.close(); sim.close()


def test_step_while_another_stream_keeps_the_sms_busy():
    """The sort stage waits on other threads (sort.cl:36-39) and is launched cooperatively: with a second stream
    saturating the SMs the step still finishes (or reports error 2) -- it never hangs (pytest timeout = failure)."""
    import torch
    n = 200_000
    a = gen(U.PlummerUniverseGenerator(17), n)
    ref, _ = parity.make_pair(a, counting=False)
    ref.step(4)
    want = ref.readBuffer("posX", n)
    ref.close()
    sim, _ = parity.make_pair(a, counting=False)
    side = torch.cuda.Stream()
    x = torch.randn(8192, 8192, device="cuda")
    with torch.cuda.stream(side):
        for _ in range(60):   # a few hundred ms of full-device GEMMs
            x = torch.tanh(x @ x) * 0.01
    try:
        sim.step(4)
        assert np.array_equal(sim.readBuffer("posX", n).view(np.uint32), want.view(np.uint32))
    except BhError as e:
        assert e.code == 2
    torch.cuda.synchronize()
    sim.close()


def test_full_size_1m_against_the_oracle():
    """BASELINE configs[1] (Plummer 2^20, theta = 0.5): every stage of two full steps against the oracle."""
    n = 1 << 20
    a = gen(U.PlummerUniverseGenerator(42), n)
    sim, orc = parity.make_pair(a, counting=True)
    for _ in range(2):
        parity.sync_oracle_from_gpu(sim, orc)
        parity.check_full_step(sim, orc)
    st = sim.stats()
    assert 0.4 < st["cells_used"] / n < 0.6 and 10 <= st["max_depth"] <= 40
    assert 2000 < st["interactions"] / n < 4000 and st["deep_walk"] == 0
    print("walk_spills at 2^20:", st["walk_spills"])
    sim.close()


def test_shared_stack_walk_kernel():
    """The second implementation of the force walk (one shared stack per warp; used for 32-wide votes), forced on for 16-wide votes."""
    a = gen(U.PlummerUniverseGenerator(9), 100_000)
    sim, orc = parity.make_pair(a)
    sim.setForceDeepWalk(True)
    for _ in range(2):
        parity.sync_oracle_from_gpu(sim, orc)
        parity.check_full_step(sim, orc)
    assert sim.stats()["deep_walk"] == 1
    sim.setForceDeepWalk(False)
    parity.sync_oracle_from_gpu(sim, orc)
    parity.check_full_step(sim, orc)
    assert sim.stats()["deep_walk"] == 0
    sim.close()


def test_very_deep_tree():
    """Pairs of bodies 1e-9 apart near the origin of a box of size ~10: chains of single-child cells 30+ levels
    deep (MAXDEPTH is 64, calculateforce.cl:12).  Whichever walk kernel ends up doing the work, the result is the oracle's."""
    n = 4096
    a = gen(U.PlummerUniverseGenerator(21), n)
    rng = np.random.default_rng(5)
    for k in range(32):
        c = (rng.random(3) - 0.5) * 1e-3
        for ax in range(3):
            a[ax][2 * k] = np.float32(c[ax])
            a[ax][2 * k + 1] = np.float32(c[ax]) + np.float32((k + 1) * 1e-9)
    assert np.unique(np.stack(a[:3], 1).view(np.uint32), axis=0).shape[0] == n
    sim, orc = parity.make_pair(a)
    parity.check_full_step(sim, orc)
    assert sim.scalar("maxDepth") >= 30
    sim.close()


def test_physical_order_is_invisible():
    """Bodies are stored in tree order internally; every logical buffer comes back in the host's numbering, before
    and after the reordering pass, and storage mode 0 (bodies stay in upload order) gives the same bits."""
    n = 30000
    a = gen(U.TwoDiskGalaxiesGenerator(3, 4), n)
    outs = []
    for mode in (1, 0):
        sim, orc = parity.make_pair(a, counting=False)
        sim.setInsertionOrder(mode)
        sim.step(3)
        assert orc.step(3) == 0
        got = {k: sim.readBuffer(k, n) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "mass", "sorted")}
        # masses never move; the body ids in sorted[] are a permutation and equal to the oracle's
        assert np.array_equal(got["mass"].view(np.uint32), a[6].view(np.uint32))
        assert np.array_equal(got["sorted"], orc.sorted[:n])
        for k in ("posX", "posY", "posZ"):
            assert np.allclose(got[k], orc.buf[k][:n], rtol=0, atol=2e-6), k
        # the child array names bodies by the host's ids: canonical structure equals the oracle's tree of step 3
        go, gc = __import__("oracle").canonicalize(sim.readBuffer("child"), n, orc.m)
        oo, oc = __import__("oracle").canonicalize(orc.child, n, orc.m)
        assert np.array_equal(gc, oc)
        # partial reads
        assert np.array_equal(sim.readBuffer("posX", 100).view(np.uint32), got["posX"][:100].view(np.uint32))
        outs.append(got)
        sim.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k].view(np.uint32), outs[1][k].view(np.uint32)), k


def test_reupload_keeps_tree_order_and_resets_the_tree_buffers():
    """An upload over a stepped state stores body i in the slot of the old body i (tree order of the last step: the
    tree stages keep their locality when the host sends the bodies every step) and clears the cells builds have used
    since the last reset.  Neither is visible: the buffers are those of a freshly created simulation, stage by stage
    against the oracle and bitwise against a fresh simulation after further steps."""
    n = 60_000
    a = gen(U.PlummerUniverseGenerator(77), n)
    sim, _ = parity.make_pair(a)
    sim.step(3)
    moved = [sim.readBuffer(k, n).copy() for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "mass")]
    b = gen(U.TwoDiskGalaxiesGenerator(8, 9), n)   # an unrelated universe goes through the same placement
    for arrays in (moved, b, a):
        sim.upload(*arrays)
        for k, w in zip(("posX", "posY", "posZ", "velX", "velY", "velZ", "mass"), arrays):   # host numbering, as uploaded
            assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), np.asarray(w, dtype=np.float32).view(np.uint32)), k
        for k in ("child", "start", "bodyCount", "sorted", "accX"):   # GPUBH:155-179: everything zero after a reset
            assert not sim.readBuffer(k).any(), k
        for k in ("posX", "posY", "posZ", "mass"):                  # ... cells included
            assert not sim.readBuffer(k)[n:].any(), k
        assert sim.scalar("step") == -1 and sim.scalar("bottom") == 0 and sim.scalar("maxDepth") == 1
        fresh, orc = parity.make_pair(arrays)
        parity.check_full_step(sim, orc)       # every stage against the oracle, on the re-uploaded simulation
        fresh.step(1)
        sim.step(2); fresh.step(2)
        for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "sorted"):   # (cell numbers are a race)
            assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), fresh.readBuffer(k, n).view(np.uint32)), k
        assert not sim.readBuffer("child")[: 8 * (n + n // 4)].any()   # rows no build has ever allocated stay zero
        fresh.close()
    # a diagnostic tree (tree stages without a step) on a fresh upload also dirties cells that the next reset must clear
    sim.upload(*a)
    sim.diagnostics(2)
    sim.upload(*b)
    for k in ("child", "start", "bodyCount"):
        assert not sim.readBuffer(k).any(), k
    sim.close()


def test_device_diagnostics_match_host_energy():
    """bh_diagnostics (printEnergy / printImpulse on the device) against the oracle's double-precision O(N^2) sum."""
    import oracle
    n = 20000
    a = gen(U.PlummerUniverseGenerator(8), n)
    sim, _ = parity.make_pair(a, counting=False)
    sim.step(5)
    d = sim.diagnostics()
    g = [sim.readBuffer(k, n) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "mass")]
    ek, ep = oracle.energy(*g)
    assert abs(d["ekin"] - ek) <= 1e-9 * abs(ek)
    assert abs(d["epot"] - ep) <= 2e-6 * abs(ep)
    m = g[6].astype(np.float64)
    for k, v in zip(("px", "py", "pz"), g[3:6]):
        assert abs(d[k] - float((m * v).sum())) < 1e-9
    assert abs(d["mass"] - float(m.sum())) < 1e-12
    sim.close()


def test_tree_potential_diagnostic():
    """bh_diagnostics mode 2: the potential through the tree (what makes an energy-drift check affordable at 10^7 bodies)
    agrees with the exact pair sum to the tree's accuracy and leaves the simulation state untouched."""
    n = 30000
    a = gen(U.PlummerUniverseGenerator(18), n)
    sim, _ = parity.make_pair(a, counting=False)
    twin, _ = parity.make_pair(a, counting=False)
    sim.step(3); twin.step(3)
    exact = sim.diagnostics(1)
    before = {k: sim.readBuffer(k, n) for k in ("posX", "posZ", "velY", "accZ", "mass")}   # (the tree buffers, sorted[] included, are rebuilt)
    step_before = sim.scalar("step")
    tree = sim.diagnostics(2)
    assert abs(tree["epot"] - exact["epot"]) <= 3e-3 * abs(exact["epot"]), (tree["epot"], exact["epot"])
    assert abs(tree["ekin"] - exact["ekin"]) <= 1e-12 * exact["ekin"] and sim.scalar("step") == step_before
    for k, v in before.items():
        assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), v.view(np.uint32)), k
    sim.step(2); twin.step(2)     # and the simulation goes on exactly as if nobody had looked
    for k in ("posX", "velY", "accZ", "sorted"):
        assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), twin.readBuffer(k, n).view(np.uint32)), k
    # a million bodies: a few milliseconds instead of an O(N^2) sum; sanity against the virial expectation of the Plummer model
    big = GPUBarnesHutNBodySimulation(Mode.DEFAULT, 1 << 20, None)
    big.init(None)
    big.generateOnDevice("plummer", 5)
    d = big.diagnostics(2)
    assert -0.7 < d["epot"] < -0.3 and 0.35 < 2 * d["ekin"] / abs(d["epot"]) < 1.3, d
    big.close(); sim.close(); twin.close()


def test_native_universe_file_upload(tmp_path):
    """bh_upload_universe_file == SerializedUniverseGenerator + loadBuffers, including the size check."""
    n = 3000
    a = gen(U.PlummerUniverseGenerator(4), n)
    path = tmp_path / "p.universe"
    U.write_universe(path, *a)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.RandomCubicUniverseGenerator(6.0, 1))
    sim.init(None)
    sim.uploadUniverseFile(path)
    for k, src in zip(("posX", "posY", "posZ", "velX", "velY", "velZ", "mass"), a):
        assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), src.view(np.uint32))
    assert sim.scalar("step") == -1
    sim.close()
    other = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n + 16, U.RandomCubicUniverseGenerator(6.0, 1))
    other.init(None)
    with pytest.raises(BhError):  # SerializedUniverseGenerator.java:41-42
        other.uploadUniverseFile(path)
    other.close()


def test_full_size_properties_10m():
    """The bench workload (BASELINE configs[2], Plummer 10^7): properties that do not need the oracle on all bodies,
    plus an oracle check of one vote-group-aligned sample of the force walk."""
    import oracle
    n = 10_000_000
    a = gen(U.PlummerUniverseGenerator(43), n)
    sim, _ = parity.make_pair(a, counting=True)
    sim.boundingBox(); sim.buildTree(); sim.summarizeTree(); sim.sort()
    m = sim.numberOfNodes
    srt = sim.readBuffer("sorted", n)
    assert np.array_equal(np.bincount(srt, minlength=n), np.ones(n, dtype=np.int64))  # a permutation
    assert sim.readBuffer("bodyCount")[m] == n
    # the oracle on the same input: whole tree bit-exact (sequential CPU build of 10^7 bodies, ~10 s), force on a sample
    orc = oracle.OracleSim(n, *a)
    orc.bounding_box(); assert orc.build_tree() == 0; orc.summarize(); orc.sort()
    assert sim.scalar("bottom") == orc.bottom[0] and sim.scalar("maxDepth") == orc.maxDepth[0]
    assert np.array_equal(srt, orc.sorted[:n])
    root_g = [sim.readBuffer(k)[m] for k in ("posX", "posY", "posZ", "mass")]
    root_o = [orc.buf[k][m] for k in ("posX", "posY", "posZ", "mass")]
    assert np.array_equal(np.array(root_g, np.float32).view(np.uint32), np.array(root_o, np.float32).view(np.uint32))
    sim.calculateForce()
    first, count = 4_000_000, 16 * 2048
    assert orc.calculate_force_range(first, count) == 0
    idx = srt[first:first + count]
    ga = np.stack([sim.readBuffer(k, n)[idx] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    oa = np.stack([orc.buf[k][idx] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    err = np.linalg.norm(ga - oa, axis=1) / np.linalg.norm(oa, axis=1)
    assert err.max() <= parity.ACC_RTOL
    st = sim.stats()
    assert 3000 < st["interactions"] / n < 3600
    assert st["walk_spills"] > 0   # a 22-level tree overflows some groups' shared-memory stacks: the spill path is part of this parity
    sim.close()


def test_argument_errors_and_async_api():
    import ctypes as C
    from gpu_nbody_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.bh_create(C.byref(h), 0, 0.5, 0.0025, 0.025, 16, 0) == -2          # nbodies < 1
    assert lib.bh_create(C.byref(h), 64, 0.5, 0.0025, 0.025, 8, 0) == -2          # vote width
    assert lib.bh_create(C.byref(h), 64, 0.5, 0.0025, 0.025, 16, 99) == -2        # device out of range
    a = gen(U.PlummerUniverseGenerator(1), 4096)
    sim, _ = parity.make_pair(a, counting=False)
    buf = np.zeros(8, np.float32)
    assert lib.bh_read(sim.handle, 99, buf.ctypes.data, 1) == -2
    assert lib.bh_read(sim.handle, 0, buf.ctypes.data, 10 ** 9) == -2
    assert lib.bh_calculate_force_slice(sim.handle, 8, 16) == -2                   # not a multiple of the vote width
    assert b"vote_width" in lib.bh_last_error(sim.handle)
    # stages out of order are refused (the tree buffers are not initialised until the stages before have run)
    fresh, _ = parity.make_pair(a, counting=False)
    for call in (lib.bh_build_tree, lib.bh_summarize, lib.bh_sort, lib.bh_calculate_force):
        assert call(fresh.handle) == -2 and b"before the stages" in lib.bh_last_error(fresh.handle)
    assert lib.bh_calculate_force_slice(fresh.handle, 0, 16) == -2
    assert lib.bh_bounding_box(fresh.handle) == 0 and lib.bh_summarize(fresh.handle) == -2 and lib.bh_build_tree(fresh.handle) == 0
    assert lib.bh_integrate(fresh.handle) == 0   # integrate needs nothing (integrate.cl only reads pos / vel / acc)
    fresh.upload(*a)
    assert lib.bh_build_tree(fresh.handle) == -2   # an upload starts over
    fresh.step(1)
    assert lib.bh_calculate_force(fresh.handle) == 0   # a stale tree is walked, as the reference would
    assert lib.bh_sort(fresh.handle) == -2             # ... but its body ids are stale after the reordering: no sort without a rebuild
    fresh.close()
    # async step + check == step
    ref, _ = parity.make_pair(a, counting=False)
    ref.step(2)
    assert lib.bh_step_async(sim.handle, 2) == 0 and lib.bh_check(sim.handle) == 0
    for k in ("posX", "velY", "accZ", "sorted"):
        assert np.array_equal(sim.readBuffer(k, 4096).view(np.uint32), ref.readBuffer(k, 4096).view(np.uint32))
    sim.close(); ref.close()


def test_device_generators():
    """bh_generate_universe: seeded, distribution of the reference generators, and a valid input for the step."""
    n = 200_000
    outs = []
    for _ in range(2):
        sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
        sim.init(None)
        sim.generateOnDevice("plummer", 7)
        outs.append([sim.readBuffer(k, n) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "mass")])
        if len(outs) == 2:
            sim.step(2)
            assert sim.scalar("error") == 0 and sim.readBuffer("bodyCount")[sim.numberOfNodes] == n
        sim.close()
    for a, b in zip(*outs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))  # same seed, same bytes
    x, y, z, vx, vy, vz, m = outs[0]
    r = np.sqrt(x.astype(np.float64) ** 2 + y ** 2 + z ** 2)
    assert abs(np.median(r) / (3 * np.pi / 16) - 1.3048) < 0.02     # Plummer half-mass radius
    assert np.all(m == np.float32(1.0 / n)) and np.unique(np.stack([x, y, z], 1).view(np.uint32), axis=0).shape[0] == n
    # virial-ish: 2T/|W| of the reference's Plummer model in its units is ~1
    ek = 0.5 * float((m.astype(np.float64) * (vx.astype(np.float64) ** 2 + vy ** 2 + vz ** 2)).sum())
    assert 0.1 < ek < 0.4

    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
    sim.init(None)
    sim.generateOnDevice("cubic", 3, 6.0)
    c = np.stack([sim.readBuffer(k, n) for k in ("posX", "posY", "posZ")], 1)
    assert c.min() >= -3.0 and c.max() <= 3.0 and abs(float(c.mean())) < 0.02 and abs(float(c.std()) - 6 / np.sqrt(12)) < 0.01
    assert not sim.readBuffer("velX", n).any()
    sim.generateOnDevice("disk", 5, 3.5, 1.0, 1.0)
    d = [sim.readBuffer(k, n) for k in ("posX", "posY", "posZ", "velX", "velY", "mass")]
    assert d[5][0] == 1.0 and d[0][0] == 0.0 and np.all(d[5][1:] == np.float32(1.0 / n))
    rr = np.sqrt(d[0][1:].astype(np.float64) ** 2 + d[1][1:] ** 2)
    assert rr.min() >= 0.05 - 1e-6 and rr.max() <= 3.55 + 1e-6 and np.abs(d[2]).max() <= 1 / 16 + 1e-6
    # circular orbits: v perpendicular to r, |v| = sqrt((M+m)/r)
    dot = d[0][1:] * d[3][1:] + d[1][1:] * d[4][1:]
    assert np.abs(dot).max() < 1e-3
    sim.step(1)
    assert sim.scalar("error") == 0
    sim.close()


def test_step_on_the_default_stream_and_on_a_torch_stream():
    """bh_set_stream with CUDA's default stream (NULL: cannot be graph-captured) and with a torch side stream."""
    import torch
    n = 8192
    a = gen(U.PlummerUniverseGenerator(6), n)
    ref, _ = parity.make_pair(a, counting=False)
    ref.step(4)
    want = [ref.readBuffer(k, n) for k in ("posX", "velY", "sorted")]
    ref.close()
    side = torch.cuda.Stream()
    for handle in (0, side.cuda_stream):
        sim, _ = parity.make_pair(a, counting=False)
        sim.setStream(handle)
        sim.step(4)
        for w, k in zip(want, ("posX", "velY", "sorted")):
            assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), w.view(np.uint32))
        sim.close()


def test_upload_from_device_memory():
    """bh_upload_device: inputs already resident in HBM (torch tensors) give the same state as bh_upload."""
    import torch
    n = 5000
    a = gen(U.PlummerUniverseGenerator(12), n)
    ref, _ = parity.make_pair(a, counting=False)
    ref.step(2)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
    sim.init(None)
    dev = [torch.from_numpy(x).cuda() for x in a]
    torch.cuda.synchronize()
    sim._check(sim._lib.bh_upload_device(sim.handle, *(t.data_ptr() for t in dev)))
    assert sim._lib.bh_use_private_stream(sim.handle) == 0
    sim.step(2)
    for k in ("posX", "posY", "velZ", "accX", "sorted"):
        assert np.array_equal(sim.readBuffer(k, n).view(np.uint32), ref.readBuffer(k, n).view(np.uint32))
    sim.close(); ref.close()


def test_uniform_1e8_sample_against_the_oracle():
    """BASELINE configs[3] size (uniform cube, 10^8 bodies, drawn on the device): the whole tree's integer outputs and a
    vote-group-aligned sample of 10^5 accelerations against the oracle, which gets its inputs through bh_read
    (SURVEY.md 8d, C4).  Needs ~40 GB of device and ~30 GB of host memory and a few minutes of sequential CPU tree build."""
    import psutil
    import torch
    import oracle
    n = 100_000_000
    if psutil.virtual_memory().available < 48e9 or torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs 48 GB of host and 60 GB of device memory")
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None)
    sim.init(None)
    sim.generateOnDevice("cubic", 44, 6.0)
    arrays = [sim.readBuffer(k, n) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "mass")]
    sim.boundingBox(); sim.buildTree(); sim.summarizeTree(); sim.sort()
    orc = oracle.OracleSim(n, *arrays)
    del arrays
    orc.bounding_box(); assert orc.build_tree() == 0; orc.summarize(); orc.sort()
    m = orc.m
    assert sim.scalar("bottom") == orc.bottom[0] and sim.scalar("maxDepth") == orc.maxDepth[0]
    srt = sim.readBuffer("sorted", n)
    assert np.array_equal(srt, orc.sorted[:n])
    assert sim.readBuffer("bodyCount")[m] == n
    root_g = np.array([sim.readBuffer(k)[m] for k in ("posX", "posY", "posZ", "mass")], np.float32)
    root_o = np.array([orc.buf[k][m] for k in ("posX", "posY", "posZ", "mass")], np.float32)
    assert np.array_equal(root_g.view(np.uint32), root_o.view(np.uint32))
    sim.setCounting(True)
    sim.calculateForce()
    first, count = 48_000_000, 16 * 6250   # 10^5 bodies = 6250 whole vote groups (calculateforce.cl:99-101)
    assert orc.calculate_force_range(first, count) == 0
    idx = srt[first:first + count]
    ga = np.stack([sim.readBuffer(k, n)[idx] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    oa = np.stack([orc.buf[k][idx] for k in ("accX", "accY", "accZ")], axis=1).astype(np.float64)
    err = np.linalg.norm(ga - oa, axis=1) / np.linalg.norm(oa, axis=1)
    assert err.max() <= parity.ACC_RTOL
    st = sim.stats()
    assert 1500 < st["interactions"] / n < 2200 and st["error"] == 0
    sim.close()
