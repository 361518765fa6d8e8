"""Stage times of the first step after an upload (bodies in the host's order) against the steady state (bodies in tree order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
a = U.generate_arrays(U.PlummerUniverseGenerator(43), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
sim.setProfiling(True)
for label, k in (("first step", 1), ("second step", 1), ("steps 3-6", 4)):
    sim.resetStats(); sim.step(k); st = sim.stats()
    print(label, {key: round(v / st["steps_timed"], 3) for key, v in st["stage_ms"].items()})
sim.upload(*a); sim.resetStats(); sim.step(1); st = sim.stats()
print("first step after re-upload", {key: round(v / st["steps_timed"], 3) for key, v in st["stage_ms"].items()})
