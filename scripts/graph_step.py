"""Wall-clock per step with and without the CUDA-graph replay (small universes are launch-bound)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
for n in (4096, 32768, 262144):
    a = U.generate_arrays(U.PlummerUniverseGenerator(42), n)
    res = {}
    for graph in (0, 1):
        sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
        sim.setGraph(graph)
        sim.step(20)
        t = time.perf_counter(); sim.step(500); dt = time.perf_counter() - t
        res["graph" if graph else "launches"] = round(dt / 500 * 1e6, 1)
        x = sim.readBuffer("posX", n)
        res.setdefault("chk", []).append(float(x.astype("float64").sum()))
        sim.close()
    res["bit_identical"] = res["chk"][0] == res["chk"][1]; del res["chk"]
    print(json.dumps({"n": n, "us_per_step": res}), flush=True)
