"""Cost of each part of bench.py's asynchronous e2e loop (development aid): python scripts/e2e_parts.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
a = U.generate_arrays(U.PlummerUniverseGenerator(43), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
lib, h = sim._lib, sim.handle
pinned = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in a]
dev = [p.cuda() for p in pinned]
pos4 = torch.empty((n, 4), dtype=torch.float32).pin_memory(); vel4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
dpos4 = torch.empty((n, 4), dtype=torch.float32, device="cuda"); dvel4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
ptrs = [p.data_ptr() for p in pinned]
dptrs = [p.data_ptr() for p in dev]
def loop(fn, reps=8):
    fn(); sim._check(lib.bh_wait_copies(h)); sim._check(lib.bh_check(h)); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    sim._check(lib.bh_wait_copies(h)); sim._check(lib.bh_check(h)); torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps
def step(): sim._check(lib.bh_step_async(h, 1))
def up(): sim._check(lib.bh_upload_async(h, *ptrs))
def up_dev(): sim._check(lib.bh_upload_device(h, *dptrs))
def cp(): sim._check(lib.bh_copy_vertices_async(h, pos4.data_ptr(), vel4.data_ptr()))
def cp_dev(): sim._check(lib.bh_copy_vertices_device(h, dpos4.data_ptr(), dvel4.data_ptr()))
sim.step(3)
rows = [("step_async only (state resident)", lambda: step()),
        ("upload_device + step", lambda: (up_dev(), step())),
        ("upload_device + step + copy_vertices_device", lambda: (up_dev(), step(), cp_dev())),
        ("upload_async + step", lambda: (up(), step())),
        ("upload_async + step + copy_vertices_device", lambda: (up(), step(), cp_dev())),
        ("upload_async + step + copy_vertices_async (bench e2e)", lambda: (up(), step(), cp())),
        ("step + copy_vertices_async", lambda: (step(), cp())),
        ("upload_async only", lambda: up()),
        ("upload_device only", lambda: up_dev())]
for name, fn in rows:
    print("%-56s %7.2f ms" % (name, loop(fn)))
