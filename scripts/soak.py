"""Soak test: many steps, error flag, energy drift (device diagnostics)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
def run(name, gen, n, steps, chunk):
    a = U.generate_arrays(gen, n)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
    e0 = sim.diagnostics()["etot"]; t = time.perf_counter()
    for _ in range(steps // chunk):
        sim.step(chunk)
    dt = time.perf_counter() - t
    d = sim.diagnostics(); st = sim.stats()
    print(json.dumps({"run": name, "n": n, "steps": steps, "s": round(dt, 2), "error": st["error"], "max_depth": st["max_depth"],
                      "energy_drift": (d["etot"] - e0) / abs(e0), "momentum": [d["px"], d["py"], d["pz"]]}), flush=True)
    sim.close()
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sphericaluniverse1.npz"))
n0 = d["x"].size; z = np.zeros(n0, np.float32)
run("bundled sphericaluniverse1 (cold collapse)", U.ArrayUniverseGenerator(d["x"], d["y"], d["z"], z, z, z, np.full(n0, d["mass"][0], np.float32)), n0, 3000, 100)
run("Plummer 2^20", U.PlummerUniverseGenerator(42), 1 << 20, 400, 50)
run("two disks 262144", U.TwoDiskGalaxiesGenerator(45, 46), 262144, 1000, 100)
