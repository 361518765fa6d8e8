import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
n = 10_000_000
a = U.generate_arrays(U.PlummerUniverseGenerator(42), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
sim.setProfiling(True)
out = []
for i in range(16):
    sim.resetStats(); sim.step(1); st = sim.stats()
    out.append((round(st["stage_ms"]["build_tree"], 3), round(st["stage_ms"]["summarize"], 3), round(st["stage_ms"]["sort"], 3)))
print(out)
