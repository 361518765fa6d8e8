"""Where the end-to-end step (host arrays in, vertices out) spends its time (development aid): python scripts/e2e_breakdown.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
a = U.generate_arrays(U.PlummerUniverseGenerator(43), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*a)); sim.init(None)
lib, h = sim._lib, sim.handle
pinned = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in a]
pos4 = torch.empty((n, 4), dtype=torch.float32).pin_memory(); vel4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
ptrs = [p.data_ptr() for p in pinned]
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / reps
def up_sync(): sim._check(lib.bh_upload(h, *ptrs))
def up_async(): sim._check(lib.bh_upload_async(h, *ptrs)); sim._check(lib.bh_check(h))
def step(): sim.step(1)
def down_sync(): sim._check(lib.bh_copy_vertices(h, pos4.data_ptr(), vel4.data_ptr()))
def loop_sync(): up_sync(); step(); down_sync()
def loop_async():
    sim._check(lib.bh_upload_async(h, *ptrs)); sim.step(1); sim._check(lib.bh_copy_vertices_async(h, pos4.data_ptr(), vel4.data_ptr()))
sim.step(3)
print("upload sync %.2f ms | upload async+check %.2f | step %.2f | copy_vertices sync %.2f | loop sync %.2f | loop async %.2f" % (
    t(up_sync), t(up_async), t(step), t(down_sync), t(loop_sync), t(loop_async)))
x = torch.empty(280_000_000 // 4, dtype=torch.float32).pin_memory(); d = torch.empty_like(x, device="cuda")
print("torch pinned H2D 280 MB %.2f ms, D2H %.2f ms" % (t(lambda: d.copy_(x, non_blocking=True)), t(lambda: x.copy_(d, non_blocking=True))))
# host-side timeline of the asynchronous loop: where does the calling thread block?
import collections
acc = collections.defaultdict(float)
torch.cuda.synchronize()
R = 6
for i in range(R + 1):
    t0 = time.perf_counter(); sim._check(lib.bh_upload_async(h, *ptrs)); t1 = time.perf_counter()
    sim.step(1); t2 = time.perf_counter()
    sim._check(lib.bh_copy_vertices_async(h, pos4.data_ptr(), vel4.data_ptr())); t3 = time.perf_counter()
    if i:
        acc["upload_async call"] += t1 - t0; acc["step call (blocks)"] += t2 - t1; acc["copy_async call"] += t3 - t2
sim._check(lib.bh_wait_copies(h))
print({k: round(1e3 * v / R, 2) for k, v in acc.items()})
sim.setProfiling(True); sim.resetStats()
for i in range(R):
    sim._check(lib.bh_upload_async(h, *ptrs)); sim.step(1); sim._check(lib.bh_copy_vertices_async(h, pos4.data_ptr(), vel4.data_ptr()))
st = sim.stats(); print("stage ms inside the loop", {k: round(v / st["steps_timed"], 3) for k, v in st["stage_ms"].items()})
