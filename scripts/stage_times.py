"""Per-stage CUDA-event times of bh_step at a few sizes (development aid)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U

def run(n, gen, steps=5, warm=3, order=1, vote=16, variant=2):
    t0 = time.time()
    a = U.generate_arrays(gen, n)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a), vote_width=vote)
    sim.init(None)
    sim.setInsertionOrder(order)
    sim.setCounting(True); sim.step(1); st = sim.stats(); sim.setCounting(False)
    inter, opens = st["interactions"], st["opens"]
    sim.step(warm - 1)
    sim.setProfiling(True); sim.resetStats()
    t1 = time.time(); sim.step(steps); wall = time.time() - t1
    st = sim.stats()
    ms = {k: v / st["steps_timed"] for k, v in st["stage_ms"].items()}
    tot = sum(ms.values())
    flops = 20 * inter + 10 * opens
    print(json.dumps({"n": n, "gen": type(gen).__name__, "order": order, "vote": vote, "variant": variant, "ms": {k: round(v, 4) for k, v in ms.items()},
                      "total_ms": round(tot, 3), "wall_ms_per_step": round(1e3 * wall / steps, 3), "body_steps_per_s": round(n / tot * 1e3),
                      "cells": st["cells_used"], "maxDepth": st["max_depth"], "I/N": round(inter / n, 1), "O/N": round(opens / n, 1),
                      "force_TFLOPs": round(flops / (ms["calculate_force"] * 1e-3) / 1e12, 2), "gen_s": round(t1 - t0, 1)}), flush=True)
    sim.close()

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32768, 1 << 20]
    for n in sizes:
        run(n, U.PlummerUniverseGenerator(42))
    run(1 << 20, U.RandomCubicUniverseGenerator(6.0, 44))
    run(1 << 20, U.PlummerUniverseGenerator(42), vote=32)
    run(1 << 20, U.PlummerUniverseGenerator(42), order=0)
