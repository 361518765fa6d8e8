"""Does a bulk PCIe transfer slow the step down, and does it matter which part of the step it overlaps?
(development aid)  python scripts/dma_interference.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
a = U.generate_arrays(U.PlummerUniverseGenerator(43), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
lib, h = sim._lib, sim.handle
main = torch.cuda.Stream(); side = torch.cuda.Stream()
sim.setStream(main.cuda_stream)
host = torch.empty(320_000_000 // 4, dtype=torch.float32).pin_memory()
dev = torch.empty(320_000_000 // 4, dtype=torch.float32, device="cuda")
hsrc = torch.empty(280_000_000 // 4, dtype=torch.float32).pin_memory()
ddst = torch.empty(280_000_000 // 4, dtype=torch.float32, device="cuda")

def staged(copy_at, direction):
    """one step as stages; the bulk copy becomes runnable at `copy_at`: 'start' of the step, after the 'tree' stages, or 'none'"""
    def go():
        if copy_at == "start": xfer(direction)
        for st in range(4): sim._check(lib.bh_stage_async(h, st))
        if copy_at == "tree": xfer(direction)
        sim._check(lib.bh_stage_async(h, 4)); sim._check(lib.bh_stage_async(h, 5))
    return go
def xfer(direction):
    ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
    with torch.cuda.stream(side):
        if direction in ("d2h", "both"): host.copy_(dev, non_blocking=True)
        if direction in ("h2d", "both"): ddst.copy_(hsrc, non_blocking=True)
def loop(fn, reps=8):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); sim._check(lib.bh_check(h))
    return 1e3 * (time.perf_counter() - t0) / reps
sim.step(3)
for at in ("none", "start", "tree"):
    for d in (("-",) if at == "none" else ("d2h", "h2d", "both")):
        print("copy runnable at %-6s %-5s %7.2f ms per step" % (at, d, loop(staged(at, d))))
