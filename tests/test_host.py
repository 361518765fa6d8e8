"""CPU: host-side logic -- .universe format, generators, node-pool rule, the C ABI's
symbols -- and the loud failure of the product path without a CUDA device."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from gpu_nbody_b200 import BhError, GPUBarnesHutNBodySimulation, Mode, _lib, universe as U
from gpu_nbody_b200.distributed import slice_bounds


def test_universe_roundtrip_and_wire_format(tmp_path):
    arrs = U.generate_arrays(U.PlummerUniverseGenerator(3), 257)
    p = tmp_path / "u.universe"
    U.write_universe(p, *arrs)
    raw = p.read_bytes()
    # 93 + 28 n bytes, Java stream magic, writeInt block, float[] class descriptor
    assert len(raw) == 93 + 28 * 257
    assert raw[:4] == b"\xac\xed\x00\x05" and raw[4:6] == b"\x77\x04" and struct.unpack(">i", raw[6:10])[0] == 257
    assert raw[10:15] == b"\x75\x72\x00\x02[" and raw[15:16] == b"F"
    n, back = U.read_universe(p)
    assert n == 257
    for a, b in zip(arrs, back):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    gen = U.SerializedUniverseGenerator(p)
    with pytest.raises(RuntimeError):  # SerializedUniverseGenerator.java:41-42
        gen.generate(0, 100, *[np.zeros(100, np.float32) for _ in range(7)])


def test_bundled_fixture_matches_appendix_b():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "sphericaluniverse1.npz"))
    r = np.sqrt(d["x"].astype(np.float64) ** 2 + d["y"].astype(np.float64) ** 2 + d["z"].astype(np.float64) ** 2)
    assert d["x"].size == 32768 and np.all(np.abs(r - 1) < 1e-6) and d["mass"][0] == np.float32(2.0 ** -15)


@pytest.mark.parametrize("gen,n", [(U.PlummerUniverseGenerator(1), 4096), (U.RandomCubicUniverseGenerator(6.0, 2), 4096),
                                   (U.TwoDiskGalaxiesGenerator(3, 4), 4096), (U.SphericalUniverseGenerator(5), 4096),
                                   (U.MonteCarloSphericalUniverseGenerator(6), 4096), (U.LonLatSphericalUniverseGenerator(7), 4096)])
def test_generators_are_seeded_and_duplicate_free(gen, n):
    a = U.generate_arrays(gen, n)
    b = U.generate_arrays(gen, n)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    pos = np.stack(a[:3], axis=1)
    assert np.unique(pos.view(np.uint32), axis=0).shape[0] == n
    assert np.all(a[6] > 0)


def test_plummer_distribution():
    x, y, z, vx, vy, vz, m = U.generate_arrays(U.PlummerUniverseGenerator(42), 200000)
    r = np.sqrt(x.astype(np.float64) ** 2 + y ** 2 + z ** 2)
    # half-mass radius of a Plummer sphere = 1.3048 a, a = 3 pi / 16 (PlummerUniverseGenerator.java:8)
    assert abs(np.median(r) / (3 * np.pi / 16) - 1.3048) < 0.02


def test_number_of_nodes_rule():
    lib = _lib.load()
    for n, m in [(1, 16384), (4096, 16384), (8192, 16384), (8193, 16400), (32768, 65536), (1000003, 2000016), (10_000_000, 20_000_000)]:
        assert lib.bh_number_of_nodes(n) == m  # GPUBH:219-227
    import oracle
    for n in (1, 5000, 8200, 123457):
        assert lib.bh_number_of_nodes(n) == oracle.number_of_nodes(n)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(lib._protos), set(declared) ^ set(lib._protos)
    assert lib.bh_abi_version() == 2


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, 64, U.PlummerUniverseGenerator(1))
    with pytest.raises(BhError) as ei:
        sim.init(None)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "gpu_nbody_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "bh_oracle" not in text, f


def test_slice_bounds():
    for n, p in [(10_000_000, 8), (1 << 20, 4), (1000, 3), (31, 2), (64, 8), (100, 8)]:
        chunk, b = slice_bounds(n, p)
        assert chunk % 32 == 0 and chunk * p >= n and chunk * p <= n + 2048
        assert sum(c for _, c in b) == n and all(f % 32 == 0 for f, _ in b)
        nonempty = [(f, c) for f, c in b if c]
        assert nonempty[0][0] == 0 and all(nonempty[i][0] == nonempty[i - 1][0] + nonempty[i - 1][1] for i in range(1, len(nonempty)))


def test_native_universe_reader_header(tmp_path):
    """bh_universe_file_bodies parses the Java stream header written by write_universe (no GPU needed)."""
    lib = _lib.load()
    arrs = U.generate_arrays(U.PlummerUniverseGenerator(3), 77)
    p = tmp_path / "u.universe"
    U.write_universe(p, *arrs)
    n = C.c_int32()
    assert lib.bh_universe_file_bodies(str(p).encode(), C.byref(n)) == 0 and n.value == 77
    bad = tmp_path / "bad.universe"
    bad.write_bytes(b"\xac\xed\x00\x05\x73\x72" + b"\0" * 64)  # the legacy layout (writeObject(Integer) first)
    assert lib.bh_universe_file_bodies(str(bad).encode(), C.byref(n)) == -2


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "universes")), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("name", ["sphericaluniverse1", "montecarlouniverse1"])
def test_readers_on_the_reference_s_own_universe_files(name):
    """universes/*.universe as shipped (UniverseSerializer.java:25-34): the Python reader and the native reader
    (bh_read_universe_file, the loader behind bh_upload_universe_file) reproduce the committed fixtures bit for bit
    (tests/golden/make_bundled_npz.py made them)."""
    path = os.path.join(REFERENCE, "universes", name + ".universe")
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    n, arrs = U.read_universe(path)
    lib = _lib.load()
    cnt = C.c_int32()
    assert lib.bh_universe_file_bodies(path.encode(), C.byref(cnt)) == 0 and cnt.value == n == 32768
    nat = [np.full(n, np.nan, np.float32) for _ in range(7)]
    assert lib.bh_read_universe_file(path.encode(), n, *(a.ctypes.data for a in nat)) == 0
    assert lib.bh_read_universe_file(path.encode(), n - 1, *(a.ctypes.data for a in nat)) == -2  # no room
    for got in (arrs, nat):
        for k, a in zip("xyz", got[:3]):
            assert np.array_equal(a.view(np.uint32), fix[k].view(np.uint32)), k
        assert not got[3].any() and not got[4].any() and not got[5].any()
        assert np.all(got[6] == fix["mass"][0])
    # the legacy layout at the repository root (writeObject(Integer) first) is rejected, as by the current Java loader
    legacy = os.path.join(REFERENCE, "sphericaluniverse1.universe")
    if os.path.exists(legacy):
        assert lib.bh_universe_file_bodies(legacy.encode(), C.byref(cnt)) == -2
        with pytest.raises(ValueError):
            U.read_universe(legacy)


def test_native_reader_rejects_bad_headers(tmp_path):
    """A negative or absurd body count in the header must come back as an error code, not as a C++ exception."""
    lib = _lib.load()
    n = C.c_int32()
    for count in (-1, 0):
        p = tmp_path / ("bad%d.universe" % count)
        p.write_bytes(b"\xac\xed\x00\x05\x77\x04" + struct.pack(">i", count) + b"\x75" + b"\0" * 40)
        assert lib.bh_universe_file_bodies(str(p).encode(), C.byref(n)) == -2
    p = tmp_path / "huge.universe"
    p.write_bytes(b"\xac\xed\x00\x05\x77\x04" + struct.pack(">i", 2 ** 31 - 1) + b"\x75" + b"\0" * 40)
    buf = [np.zeros(4, np.float32) for _ in range(7)]
    assert lib.bh_read_universe_file(str(p).encode(), 4, *(a.ctypes.data for a in buf)) == -2
