"""GPU, >= 2 devices: the NCCL multi-GPU step (gpu_nbody_b200.distributed) leaves every rank with
state bit-identical to the single-GPU step (slices are aligned to vote groups, the tree is replicated)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
from gpu_nbody_b200.distributed import CudaSliceEngine, DistributedBarnesHutSimulation
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, steps = %(n)d, %(steps)d
arrays = U.generate_arrays(U.PlummerUniverseGenerator(123), n)
sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*arrays), device=local)
sim.init(None)
dsim = DistributedBarnesHutSimulation(CudaSliceEngine(sim, p2p=%(p2p)s), rank, world)
dsim.step(steps)
np.savez(os.path.join(%(out)r, "rank%%d.npz" %% rank), **{k: sim.readBuffer(k, n) for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "sorted")})
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("n", [100000, 4097])
def test_ranks_equal_single_gpu(tmp_path, n, p2p):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    steps = 3
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "n": n, "steps": steps, "out": str(tmp_path), "p2p": p2p})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), str(script)]
    subprocess.run(cmd, check=True, timeout=600)
    from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
    arrays = U.generate_arrays(U.PlummerUniverseGenerator(123), n)
    sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*arrays))
    sim.init(None)
    sim.step(steps)
    for r in range(world):
        got = np.load(tmp_path / ("rank%d.npz" % r))
        for k in got.files:
            assert np.array_equal(got[k].view(np.uint32), sim.readBuffer(k, n).view(np.uint32)), (r, k)
    sim.close()
