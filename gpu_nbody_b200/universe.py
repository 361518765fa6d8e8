"""Universe generators and the ``.universe`` file format.

Host-side mirror of the reference's ``ch.fhnw.woipv.nbody.simulation.universe``
package (paths relative to /root/reference):

* ``UniverseGenerator.generate(offset, nbodies, x, y, z, vx, vy, vz, mass)``
  (universe/UniverseGenerator.java:22) fills caller-owned SoA float arrays; the
  classes here keep that signature.
* The reference generators draw from the unseeded ``Math.random()``; these take a
  seed (``numpy.random.default_rng``) and reproduce the *distributions*, which is
  what benchmarks and parity tests need (same bytes into oracle and CUDA path).
* ``read_universe`` / ``write_universe`` keep the Java ``ObjectOutputStream`` wire
  format of universe/serialize/UniverseSerializer.java:25-34 byte for byte, so
  files written here load in ``SerializedUniverseGenerator`` and vice versa.
"""
from __future__ import annotations

import math
import struct

import numpy as np

__all__ = [
    "UniverseGenerator", "PlummerUniverseGenerator", "RandomCubicUniverseGenerator",
    "SphericalUniverseGenerator", "MonteCarloSphericalUniverseGenerator",
    "LonLatSphericalUniverseGenerator", "RotatingDiskGalaxyGenerator", "TwoDiskGalaxiesGenerator",
    "SerializedUniverseGenerator", "ArrayUniverseGenerator",
    "TwoBodyUniverse", "EightBodyUniverse", "BigTreeUniverse",
    "read_universe", "write_universe", "generate_arrays",
]


class UniverseGenerator:
    """universe/UniverseGenerator.java:22."""

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        raise NotImplementedError


def generate_arrays(generator: UniverseGenerator, nbodies: int):
    """Allocate seven float32 arrays of length ``nbodies`` and fill them."""
    arrs = [np.zeros(nbodies, dtype=np.float32) for _ in range(7)]
    generator.generate(0, nbodies, *arrs)
    return arrs


def _unit_vectors(rng, n):
    """Rejection sampling of a point in the unit ball, normalised (the do/while
    loops of PlummerUniverseGenerator.java:15-20,31-36), vectorised."""
    out = np.empty((n, 3), dtype=np.float64)
    todo = np.arange(n)
    while todo.size:
        v = rng.random((todo.size, 3)) * 2.0 - 1.0
        sq = (v * v).sum(axis=1)
        ok = (sq <= 1.0) & (sq > 0.0)
        out[todo[ok]] = v[ok] / np.sqrt(sq[ok])[:, None]
        todo = todo[~ok]
    return out


def _dedupe(x, y, z, rng, scale):
    """Identical fp32 positions make the reference subdivide until the cell pool
    is exhausted (buildtree.cl:106-119); nudge exact duplicates apart."""
    for _ in range(8):
        key = np.stack([x, y, z], axis=1).view(np.uint32)
        _, first = np.unique(key, axis=0, return_index=True)
        if first.size == x.size:
            return
        dup = np.ones(x.size, dtype=bool)
        dup[first] = False
        k = int(dup.sum())
        x[dup] += ((rng.random(k) - 0.5) * 1e-3 * scale).astype(np.float32)
        y[dup] += ((rng.random(k) - 0.5) * 1e-3 * scale).astype(np.float32)
        z[dup] += ((rng.random(k) - 0.5) * 1e-3 * scale).astype(np.float32)
    raise RuntimeError("could not de-duplicate positions")


class PlummerUniverseGenerator(UniverseGenerator):
    """universe/PlummerUniverseGenerator.java:8-41 (ignores ``offset`` like the reference)."""

    def __init__(self, seed: int = 42):
        self.seed = seed

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        rng = np.random.default_rng(self.seed)
        n = nbodies
        rsc = (3 * math.pi) / 16
        vsc = math.sqrt(1.0 / rsc)
        bodiesMass[:n] = np.float32(1.0 / n)
        r = 1.0 / np.sqrt(np.power(rng.random(n) * 0.999, -2.0 / 3.0) - 1)
        d = _unit_vectors(rng, n)
        scale = rsc * r
        bodiesX[:n] = (d[:, 0] * scale).astype(np.float32)
        bodiesY[:n] = (d[:, 1] * scale).astype(np.float32)
        bodiesZ[:n] = (d[:, 2] * scale).astype(np.float32)
        # speed: rejection sample x in [0,1), y in [0,0.1) with y <= x^2 (1-x^2)^3.5  (:25-28)
        xs = np.empty(n, dtype=np.float64)
        todo = np.arange(n)
        while todo.size:
            x = rng.random(todo.size)
            y = rng.random(todo.size) * 0.1
            ok = y <= x * x * np.power(1 - x * x, 3.5)
            xs[todo[ok]] = x[ok]
            todo = todo[~ok]
        v = xs * np.sqrt(2.0 / np.sqrt(1 + r * r))
        d = _unit_vectors(rng, n)
        scale = vsc * v
        velX[:n] = (d[:, 0] * scale).astype(np.float32)
        velY[:n] = (d[:, 1] * scale).astype(np.float32)
        velZ[:n] = (d[:, 2] * scale).astype(np.float32)
        _dedupe(bodiesX[:n], bodiesY[:n], bodiesZ[:n], rng, 1e-3)


class RandomCubicUniverseGenerator(UniverseGenerator):
    """universe/RandomCubicUniverseGenerator.java:13-17: (U-0.5)*range per axis, v = 0, m = 1/n."""

    def __init__(self, range_: float, seed: int = 44):
        self.range = float(range_)
        self.seed = seed

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, mass):
        rng = np.random.default_rng(self.seed)
        sl = slice(offset, offset + nbodies)
        for arr in (bodiesX, bodiesY, bodiesZ):
            arr[sl] = ((rng.random(nbodies) - 0.5) * self.range).astype(np.float32)
        mass[sl] = np.float32(1.0) / np.float32(nbodies)
        _dedupe(bodiesX[sl], bodiesY[sl], bodiesZ[sl], rng, self.range * 1e-3)


class SphericalUniverseGenerator(UniverseGenerator):
    """universe/SphericalUniverseGenerator.java: uniform on the unit sphere surface."""

    def __init__(self, seed: int = 47):
        self.seed = seed

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        rng = np.random.default_rng(self.seed)
        n = nbodies
        omega = rng.random(n) * 2 * math.pi
        u = rng.random(n) * 2 - 1
        s = np.sqrt(1 - u * u)
        bodiesX[:n] = (s * np.cos(omega)).astype(np.float32)
        bodiesY[:n] = (s * np.sin(omega)).astype(np.float32)
        bodiesZ[:n] = u.astype(np.float32)
        bodiesMass[:n] = np.float32(1.0) / np.float32(n)
        _dedupe(bodiesX[:n], bodiesY[:n], bodiesZ[:n], rng, 1e-3)


class MonteCarloSphericalUniverseGenerator(UniverseGenerator):
    """universe/MonteCarloSphericalUniverseGenerator.java: golden-angle spiral on the sphere."""

    def __init__(self, seed: int = 48):
        self.seed = seed

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        rng = np.random.default_rng(self.seed)
        n = nbodies
        rand = rng.random() * n
        o = 2 / float(n)
        increment = math.pi * (3.0 - math.sqrt(5))
        i = np.arange(n, dtype=np.float64)
        y = ((i * o) - 1) + (o / 2)
        r = np.sqrt(1 - y * y)
        phi = ((i + rand) % n) * increment
        bodiesX[:n] = (np.cos(phi) * r).astype(np.float32)
        bodiesY[:n] = y.astype(np.float32)
        bodiesZ[:n] = (np.sin(phi) * r).astype(np.float32)
        bodiesMass[:n] = np.float32(1.0) / np.float32(n)


class LonLatSphericalUniverseGenerator(UniverseGenerator):
    """universe/LonLatSphericalUniverseGenerator.java (R = 5)."""

    def __init__(self, seed: int = 49):
        self.seed = seed

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        rng = np.random.default_rng(self.seed)
        n = nbodies
        lon = rng.random(n) * 2 * math.pi
        lat = rng.random(n) * 2 * math.pi
        bodiesX[:n] = (5.0 * np.cos(lat) * np.cos(lon)).astype(np.float32)
        bodiesY[:n] = (5.0 * np.cos(lat) * np.sin(lon)).astype(np.float32)
        bodiesZ[:n] = (5.0 * np.sin(lat)).astype(np.float32)
        bodiesMass[:n] = np.float32(1.0) / np.float32(n)
        _dedupe(bodiesX[:n], bodiesY[:n], bodiesZ[:n], rng, 5e-3)


class RotatingDiskGalaxyGenerator(UniverseGenerator):
    """universe/RotatingDiskGalaxyGenerator.java:17-43.  Body 0 carries ``centerMass``
    at the origin; velZ is never written (stays as passed in)."""

    RADIUS_OFFSET = np.float32(0.05)

    def __init__(self, r: float, velocityMultiplier: float, centerMass: float, seed: int = 45):
        self.radius = np.float32(r)
        self.velocityMultiplier = np.float32(velocityMultiplier)
        self.centerMass = np.float32(centerMass)
        self.seed = seed

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        rng = np.random.default_rng(self.seed)
        n = nbodies
        bodiesMass[0] = self.centerMass
        k = n - 1
        r = (rng.random(k) * float(self.radius)).astype(np.float32) + self.RADIUS_OFFSET
        alpha = rng.random(k) * 2 * math.pi
        x = (np.cos(alpha) * r).astype(np.float32)
        y = (np.sin(alpha) * r).astype(np.float32)
        bodiesX[1:n] = x
        bodiesY[1:n] = y
        bodiesZ[1:n] = ((rng.random(k) - 0.5) / 8).astype(np.float32)
        m = np.float32(1.0) / np.float32(n)
        bodiesMass[1:n] = m
        v0 = np.sqrt((self.centerMass + m) / (r * r * r)).astype(np.float32) * self.velocityMultiplier
        velX[1:n] = y * v0
        velY[1:n] = -x * v0
        _dedupe(bodiesX[:n], bodiesY[:n], bodiesZ[:n], rng, 1e-3)


class TwoDiskGalaxiesGenerator(UniverseGenerator):
    """Two RotatingDiskGalaxyGenerator(3.5, 1, 1) disks (the parameters of
    NBodyVisualizer.java:212) on a collision course -- SURVEY.md 8(d) config C5.
    The offsets and bulk velocities are this repo's (the reference has no
    two-galaxy generator): disk A at (-4,0,0) moving +0.25 x, disk B at (4,1,0.5)
    moving -0.25 x."""

    def __init__(self, seed_a: int = 45, seed_b: int = 46):
        self.seed_a, self.seed_b = seed_a, seed_b

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        na = nbodies // 2
        nb = nbodies - na
        a = generate_arrays(RotatingDiskGalaxyGenerator(3.5, 1.0, 1.0, self.seed_a), na)
        b = generate_arrays(RotatingDiskGalaxyGenerator(3.5, 1.0, 1.0, self.seed_b), nb)
        a[0] += np.float32(-4.0); a[3] += np.float32(0.25)
        b[0] += np.float32(4.0); b[1] += np.float32(1.0); b[2] += np.float32(0.5); b[3] += np.float32(-0.25)
        for dst, pa, pb in zip((bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass), a, b):
            dst[:na] = pa
            dst[na:nbodies] = pb


class TwoBodyUniverse(UniverseGenerator):
    """universe/test/TwoBodyUniverse.java"""

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        bodiesX[0], bodiesY[0], bodiesZ[0], bodiesMass[0] = -1.1, -1, -1, 1 / 8
        bodiesX[1], bodiesY[1], bodiesZ[1], bodiesMass[1] = 1, 1.1, 1, 1 / 8


class EightBodyUniverse(UniverseGenerator):
    """universe/test/EightBodyUniverse.java"""

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        pts = [(-1.1, -1, -1), (1, -1, -1), (-1, 1, -1.1), (-1, -1, 1.1),
               (1, 1, -1), (-1, 1.1, 1), (1, -1, 1), (1, 1.1, 1)]
        for i, (x, y, z) in enumerate(pts):
            bodiesX[i], bodiesY[i], bodiesZ[i], bodiesMass[i] = x, y, z, 1 / 8


class BigTreeUniverse(UniverseGenerator):
    """universe/test/BigTreeUniverse.java (body 3 keeps mass 0, as in the reference)."""

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        bodiesX[0], bodiesY[0], bodiesZ[0], bodiesMass[0] = 0, 0, 0, 0.1
        bodiesX[1], bodiesY[1], bodiesZ[1], bodiesMass[1] = 0.00000001, 0, 0, 0.1
        bodiesX[2], bodiesY[2], bodiesZ[2], bodiesMass[2] = 10000, 10000, 10000, 0.1
        bodiesX[3], bodiesY[3], bodiesZ[3] = -10000, -10000, -10000


class ArrayUniverseGenerator(UniverseGenerator):
    """Wraps seven in-memory arrays; same size check as SerializedUniverseGenerator."""

    def __init__(self, x, y, z, vx, vy, vz, mass):
        self.arrays = [np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z, vx, vy, vz, mass)]
        self.nbodies = int(self.arrays[0].size)
        if any(a.size != self.nbodies for a in self.arrays):
            raise ValueError("universe arrays differ in length")

    def generate(self, offset, nbodies, bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass):
        if nbodies != self.nbodies:
            # SerializedUniverseGenerator.java:41-42 throws IllegalStateException
            raise RuntimeError("invalid amount of bodies for serialized universe")
        for dst, src in zip((bodiesX, bodiesY, bodiesZ, velX, velY, velZ, bodiesMass), self.arrays):
            dst[offset:offset + nbodies] = src


class SerializedUniverseGenerator(ArrayUniverseGenerator):
    """universe/serialize/SerializedUniverseGenerator.java:21-53."""

    def __init__(self, file):
        n, arrays = read_universe(file)
        super().__init__(*arrays)
        assert n == self.nbodies


# --------------------------------------------------------------------------- #
# .universe wire format
# --------------------------------------------------------------------------- #
_MAGIC = b"\xac\xed\x00\x05"
_FLOAT_ARRAY_CLASSDESC = (b"\x72\x00\x02[F" + bytes.fromhex("0b9c818922e00c42") + b"\x02\x00\x00\x78\x70")
_FLOAT_ARRAY_REF = b"\x71\x00\x7e\x00\x00"


def write_universe(path, x, y, z, vx, vy, vz, mass):
    """UniverseSerializer.java:25-34: writeInt(n) then 7 x writeObject(float[n])."""
    arrays = [np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, z, vx, vy, vz, mass)]
    n = arrays[0].size
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(b"\x77\x04" + struct.pack(">i", n))
        for i, a in enumerate(arrays):
            f.write(b"\x75")
            f.write(_FLOAT_ARRAY_CLASSDESC if i == 0 else _FLOAT_ARRAY_REF)
            f.write(struct.pack(">i", a.size))
            f.write(a.astype(">f4").tobytes())


def read_universe(path):
    """Returns ``(nbodies, [x, y, z, vx, vy, vz, mass])`` (float32, native order)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != _MAGIC:
        raise ValueError("not a Java serialization stream")
    if data[4:6] != b"\x77\x04":
        raise ValueError("unsupported .universe layout (expected writeInt block first)")
    n = struct.unpack(">i", data[6:10])[0]
    off = 10
    arrays = []
    for i in range(7):
        if data[off] != 0x75:
            raise ValueError("expected TC_ARRAY at offset %d" % off)
        off += 1
        if data[off] == 0x72:
            if data[off:off + len(_FLOAT_ARRAY_CLASSDESC)] != _FLOAT_ARRAY_CLASSDESC:
                raise ValueError("unexpected class descriptor (want float[])")
            off += len(_FLOAT_ARRAY_CLASSDESC)
        elif data[off:off + 5] == _FLOAT_ARRAY_REF:
            off += 5
        else:
            raise ValueError("unexpected array class at offset %d" % off)
        length = struct.unpack(">i", data[off:off + 4])[0]
        off += 4
        arrays.append(np.frombuffer(data, dtype=">f4", count=length, offset=off).astype(np.float32))
        off += 4 * length
    return n, arrays
