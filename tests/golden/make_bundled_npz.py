"""Converts the reference's bundled universes (universes/*.universe, Java ObjectOutputStream layout,
UniverseSerializer.java:25-34) into the small .npz fixtures the tests load on boxes without /root/reference.

    python tests/golden/make_bundled_npz.py [/root/reference]

Only x, y, z and one mass are kept: both files have zero velocities and equal masses (checked here).  tests/test_host.py re-reads the
original files -- with the Python reader and with the native one (bh_read_universe_file) -- wherever the
reference tree exists and compares them with these fixtures bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gpu_nbody_b200 import universe as U  # noqa: E402


def main(ref="/root/reference"):
    for name in ("sphericaluniverse1", "montecarlouniverse1"):
        n, (x, y, z, vx, vy, vz, mass) = U.read_universe(os.path.join(ref, "universes", name + ".universe"))
        assert n == x.size == 32768 and not vx.any() and not vy.any() and not vz.any() and np.all(mass == mass[0])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, y=y, z=z, mass=mass[:1])
        print(name, n, "bodies")


if __name__ == "__main__":
    main(*sys.argv[1:])
