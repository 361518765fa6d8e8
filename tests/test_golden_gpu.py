"""GPU: the CUDA path (through the C ABI) directly against the golden vectors produced by
the reference's own kernel sources -- no oracle in between for the integer structure."""
import numpy as np
import pytest

import goldencheck as gc
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, _lib, universe as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gc.REFERENCE_FIXTURES)
def test_cuda_matches_reference_kernels(name):
    g = gc.load(name)
    arrays = gc.inputs(name, g)
    n = int(g["n"])
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*arrays), eps2=float(g["eps2"]), dt=float(g["dt"]),
                                      theta_macro=float(g["theta_macro"]))
    sim.init(None)
    sim.step(int(g["steps"]) - 1)
    sim.boundingBox(); sim.buildTree(); sim.summarizeTree(); sim.sort(); sim.calculateForce()
    buf = {k: sim.readBuffer(k) for k in _lib.BUFFERS}
    err = gc.check_after_force(buf, g)
    assert err < 1e-5  # in practice the CUDA path is within a few ulp of the reference kernels
    sim.integrate()
    gc.check_after_integrate({k: sim.readBuffer(k) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ")}, g)
    sim.close()


def test_cuda_ten_steps_of_the_bundled_universe():
    """BASELINE configs[0]: sphericaluniverse1, theta = 0.5, 10 steps, against the reference kernels' own 10 steps."""
    g = gc.load(gc.TRAJECTORY_FIXTURE)
    arrays = gc.bundled_inputs("sphericaluniverse1")
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, 32768, U.ArrayUniverseGenerator(*arrays), eps2=float(g["eps2"]), dt=float(g["dt"]),
                                      theta_macro=float(g["theta_macro"]))
    sim.init(None)
    sim.step(10)
    err = gc.check_trajectory({k: sim.readBuffer(k) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "step", "maxDepth")}, g)
    assert err < 1e-5
    sim.close()
