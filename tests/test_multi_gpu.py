"""GPU, >= 2 devices: the multi-GPU step (gpu_nbody_b200.distributed) leaves every rank with state bit-identical to
the single-GPU step (slices are aligned to vote groups, the tree is replicated).  Three transports are covered:
the fused one (peer stores over CUDA IPC + device barrier, the whole step inside bh_step_async / a CUDA graph), the
NCCL all-gather, and the fallback from the first to the second when peer mapping fails on ONE rank only."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "sorted")

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
from gpu_nbody_b200.distributed import CudaSliceEngine, DistributedBarnesHutSimulation
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, steps, mode = %(n)d, %(steps)d, %(mode)r
arrays = U.generate_arrays(U.PlummerUniverseGenerator(123), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*arrays), device=local)
sim.init(None)
engine = CudaSliceEngine(sim, p2p=(mode != "nccl"))
if mode == "fallback" and rank == world - 1:
    # peer mapping "fails" on this rank only: the others have already mapped their peers when they learn about it
    def broken(*a, **k):
        raise RuntimeError("simulated cudaIpcOpenMemHandle failure")
    engine.connect_peers_real, engine.connect_peers = engine.connect_peers, broken
    import ctypes as C
    mine = C.create_string_buffer(64)
    sim._check(sim._lib.bh_ipc_export(sim.handle, mine))
    handles = [None] * world
    dist.all_gather_object(handles, mine.raw)   # take part in the handle exchange the other ranks are in
dsim = DistributedBarnesHutSimulation(engine, rank, world)
assert dsim.fused == (mode == "fused"), (mode, dsim.fused, dsim.peer_error)
dsim.step(1)
dsim.step(steps - 1)
out = {k: sim.readBuffer(k, n) for k in %(keys)r}
out["launches"] = np.array([sum(sim.stats()["stage_launches"].values())])
np.savez(os.path.join(%(out)r, "rank%%d.npz" %% rank), **out)
dist.barrier()
dist.destroy_process_group()
'''


def _world():
    import torch
    return min(torch.cuda.device_count(), 8)


@pytest.mark.parametrize("mode", ["fused", "nccl", "fallback"])
@pytest.mark.parametrize("n", [100000, 4097])
def test_ranks_equal_single_gpu(tmp_path, n, mode):
    world = _world()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    steps = 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "n": n, "steps": steps, "out": str(tmp_path), "mode": mode, "keys": KEYS})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), str(script)]
    subprocess.run(cmd, check=True, timeout=900)
    from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
    arrays = U.generate_arrays(U.PlummerUniverseGenerator(123), n)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*arrays))
    sim.init(None)
    sim.step(steps)
    for r in range(world):
        got = np.load(tmp_path / ("rank%d.npz" % r))
        for k in KEYS:
            assert np.array_equal(got[k].view(np.uint32), sim.readBuffer(k, n).view(np.uint32)), (r, k)
        assert got["launches"][0] > 0
    sim.close()
