"""Time the force walk of each of P equal sorted slices on ONE GPU: the load balance the multi-GPU run would see."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
from gpu_nbody_b200.distributed import slice_bounds

def run(name, gen, n, P=8):
    a = U.generate_arrays(gen, n)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a))
    sim.init(None)
    sim.setStream(torch.cuda.current_stream().cuda_stream)
    sim.step(3)
    lib, h = sim._lib, sim.handle
    for st in range(4):
        lib.bh_stage_async(h, st)
    chunk, bounds = slice_bounds(n, P)
    times = []
    for first, count in bounds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.bh_calculate_force_slice(h, first, count)  # warm
        e0.record(); lib.bh_calculate_force_slice(h, first, count); e1.record(); torch.cuda.synchronize()
        times.append(round(e0.elapsed_time(e1), 3))
    print(json.dumps({"config": name, "n": n, "slices": P, "force_ms_per_slice": times, "max_over_mean": round(max(times) / (sum(times) / P), 3)}), flush=True)
    sim.close()

if __name__ == "__main__":
    run("Plummer 10^7", U.PlummerUniverseGenerator(43), 10_000_000)
    run("two disks 4M", U.TwoDiskGalaxiesGenerator(45, 46), 4_000_000)
    run("uniform 10^7", U.RandomCubicUniverseGenerator(6.0, 44), 10_000_000)
