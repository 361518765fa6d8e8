"""BASELINE.json configs on one GPU: stage times, body-steps/s, interactions per body (development aid / profiles)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U

def run(name, gen, n, theta=0.5, steps=5, warm=3):
    a = U.generate_arrays(gen, n)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a), theta=theta)
    sim.init(None)
    sim.step(warm - 1)
    sim.setCounting(True); sim.step(1); st = sim.stats(); sim.setCounting(False)
    inter, opens = st["interactions"], st["opens"]
    sim.setProfiling(True); sim.resetStats(); sim.step(steps); st = sim.stats()
    ms = {k: v / st["steps_timed"] for k, v in st["stage_ms"].items()}
    tot = sum(ms.values())
    print(json.dumps({"config": name, "n": n, "theta": theta, "ms_per_step": round(tot, 3), "body_steps_per_s": round(n / tot * 1e3),
                      "stage_ms": {k: round(v, 4) for k, v in ms.items()}, "cells_per_body": round(st["cells_used"] / n, 3),
                      "max_depth": st["max_depth"], "I_per_body": round(inter / n, 1), "O_per_body": round(opens / n, 1),
                      "force_tflops": round((20 * inter + 10 * opens) / (ms["calculate_force"] * 1e-3) / 1e12, 2)}), flush=True)
    sim.close()

if __name__ == "__main__":
    import numpy as np
    d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sphericaluniverse1.npz"))
    n0 = d["x"].size; z = np.zeros(n0, np.float32)
    run("C1 bundled sphericaluniverse1", U.ArrayUniverseGenerator(d["x"], d["y"], d["z"], z, z, z, np.full(n0, d["mass"][0], np.float32)), n0, steps=10)
    run("C2 Plummer 2^20", U.PlummerUniverseGenerator(42), 1 << 20)
    run("C3 Plummer 10^7 (1 GPU)", U.PlummerUniverseGenerator(43), 10_000_000)
    run("C4' uniform cube 10^7 (1 GPU share of the 10^8 config)", U.RandomCubicUniverseGenerator(6.0, 44), 10_000_000)
    for th in (0.3, 0.4, 0.5, 0.6, 0.7, 0.8):
        run("C5 two disks 4M", U.TwoDiskGalaxiesGenerator(45, 46), 4_000_000, theta=th, steps=3)
