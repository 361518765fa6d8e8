"""CPU, world_size 2 over gloo: the multi-GPU step of gpu_nbody_b200.distributed with
the CPU oracle standing in for the CUDA engine gives bit-identical state to the
single-process step on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleSliceEngine:
    """The slice contract of include/bhstep.h (bh_calculate_force_slice / bh_apply_acceleration) on the oracle."""

    def __init__(self, orc):
        self.o = orc
        self.nbodies = orc.n
        self.acc_sorted = torch.zeros((orc.n + 2048, 4), dtype=torch.float32)

    def tree_stages(self):
        o = self.o
        o.bounding_box(); assert o.build_tree() == 0; o.summarize(); o.sort()
        n = o.n
        self._old = [o.buf[k][:n].copy() for k in ("accX", "accY", "accZ")]
        self._vel = [o.buf[k][:n].copy() for k in ("velX", "velY", "velZ")]

    def force_slice(self, first, count):
        o, n = self.o, self.o.n
        if count:
            step = o.buf["step"][0]
            o.buf["step"][0] = 0          # slice mode: no velocity correction inside the walk
            assert o.calculate_force_range(first, count) == 0
            o.buf["step"][0] = step
            idx = o.sorted[first:first + count]
            a = np.stack([o.buf[k][idx] for k in ("accX", "accY", "accZ")] + [np.zeros(count, np.float32)], axis=1)
            self.acc_sorted[first:first + count] = torch.from_numpy(a)
        for k, old in zip(("accX", "accY", "accZ"), self._old):
            o.buf[k][:n] = old            # accelerations are only applied after the all-gather

    def apply_and_integrate(self):
        o, n = self.o, self.o.n
        a = self.acc_sorted[:n].numpy()
        idx = o.sorted[:n]
        dt = np.float32(o.state.timestep)
        for c, (ka, kv) in enumerate(zip(("accX", "accY", "accZ"), ("velX", "velY", "velZ"))):
            new = np.empty(n, np.float32); new[idx] = a[:, c]
            if o.buf["step"][0] > 0:      # calculateforce.cl:174-179
                o.buf[kv][:n] = o.buf[kv][:n] + ((new - o.buf[ka][:n]) * dt) * np.float32(0.5)
            o.buf[ka][:n] = new
        o.integrate()

    def check(self):
        pass


def _worker(rank, world, port, n, steps, out_dir):
    sys.path.insert(0, ROOT)
    import oracle
    from gpu_nbody_b200 import universe as U
    from gpu_nbody_b200.distributed import DistributedBarnesHutSimulation
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arrays = U.generate_arrays(U.PlummerUniverseGenerator(77), n)
    eng = OracleSliceEngine(oracle.OracleSim(n, *arrays))
    dsim = DistributedBarnesHutSimulation(eng, rank, world)
    dsim.step(steps)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{k: eng.o.buf[k][:n] for k in ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ")})
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4096, 1000])
def test_two_ranks_equal_single_process(tmp_path, n):
    import oracle
    from gpu_nbody_b200 import universe as U
    steps, world = 3, 2
    port = 29600 + (os.getpid() + n) % 300
    mp.start_processes(_worker, args=(world, port, n, steps, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    single = oracle.OracleSim(n, *U.generate_arrays(U.PlummerUniverseGenerator(77), n))
    assert single.step(steps) == 0
    for r in range(world):
        got = np.load(tmp_path / ("rank%d.npz" % r))
        for k in got.files:
            assert np.array_equal(got[k].view(np.uint32), single.buf[k][:n].view(np.uint32)), (r, k)
