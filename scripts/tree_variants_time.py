import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
n = 10_000_000
a = U.generate_arrays(U.PlummerUniverseGenerator(43), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
sim.step(3); sim.setProfiling(True); sim.resetStats(); sim.step(6); st = sim.stats()
print(os.environ.get("BHSTEP_LIBRARY", "default"), {k: round(v / st["steps_timed"], 3) for k, v in st["stage_ms"].items()})
