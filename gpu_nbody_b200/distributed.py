"""Multi-GPU Barnes-Hut step: one process per GPU, torch.distributed for the plumbing.

The reference is single-device (one CLDevice, GPUBarnesHutNBodySimulation.java:120);
this is the one natural shard of its step (DESIGN.md "Multi-GPU"):

  every rank   bounding box -> build tree -> summarise -> sort   (replicated, identical trees)
  rank p       force walk for the sorted slots [first_p, first_p + count_p)
  all ranks    all-gather of the sorted-order acceleration slices (16 B per body over NVLink)
  every rank   velocity correction + integrate + reorder for all bodies (replicated, one fused pass)

Slices are equal-sized multiples of 32 sorted slots, so vote groups are exactly
those of the single-GPU run and the all-gather is in place.  Two transports:

* `p2p=True` (default in bench.py) -- the all-gather is fused into the force
  kernel: every rank's acceleration buffer is mapped into every other rank
  (CUDA IPC over NVLink) and the walk's epilogue stores each result straight
  into all of them, so the transfer overlaps the walk.  The cross-rank barrier
  before the buffer is read is a device kernel on flags in the same mapped
  allocation, so the whole sliced step is inside the library (`bh_set_slice` +
  `bh_step_async`: one CUDA graph per step, no host or NCCL call per step).
* NCCL (`all_gather_into_tensor`) between `bh_calculate_force_slice` and
  `bh_finish_async`; also the fallback when peer mapping fails on any rank.

What crosses NVLink is 16 B per body per step (float4 acceleration); positions
never travel because the finish pass is replicated and bit-deterministic.

`engine` is anything with the stage methods below (the CUDA simulation in
production; the CPU oracle in the world_size-2 gloo tests), the all-gather runs
on the engine's sorted-order acceleration buffer.
"""
from __future__ import annotations

from .simulation import GPUBarnesHutNBodySimulation

SLICE_ALIGN = 32  # a hardware warp = two 16-body vote groups


def slice_bounds(nbodies: int, world_size: int, align: int = SLICE_ALIGN):
    """Equal slices of `chunk` sorted slots (chunk a multiple of `align`); the last
    non-empty slice is cut at nbodies.  Returns (chunk, [(first, count)] per rank)."""
    chunk = -(-nbodies // world_size)
    chunk = -(-chunk // align) * align
    bounds = []
    for r in range(world_size):
        first = r * chunk
        count = max(0, min(chunk, nbodies - first))
        bounds.append((first, count) if count else (0, 0))
    return chunk, bounds


class CudaSliceEngine:
    """Adapter from GPUBarnesHutNBodySimulation to the slice contract of include/bhstep.h."""

    def __init__(self, sim: GPUBarnesHutNBodySimulation, p2p: bool = False):
        import torch
        self.p2p = p2p
        self.sim = sim
        self.lib = sim._lib
        self.h = sim.handle
        self.nbodies = sim.nbodies
        self.device = torch.device("cuda", sim.device)
        ptr = self.lib.bh_acc_sorted_device_ptr(self.h)
        cap = sim.nbodies + 2048

        class _Buf:
            __cuda_array_interface__ = {"shape": (cap, 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                        "strides": None}
        self._keep = _Buf()
        self.acc_sorted = torch.as_tensor(self._keep, device=self.device)
        assert self.acc_sorted.data_ptr() == int(ptr)
        # run on torch's current stream so that NCCL ops and kernels are ordered
        sim.setStream(torch.cuda.current_stream(self.device).cuda_stream)

    def _rc(self, rc):
        self.sim._check(rc)

    def tree_stages(self):
        for st in range(4):  # bbox, build, summarise, sort
            self._rc(self.lib.bh_stage_async(self.h, st))

    def connect_peers(self, rank, world_size, group=None):
        """Exchange the CUDA IPC handles of the acceleration buffers (host side, once)."""
        import ctypes as C
        import torch.distributed as dist
        mine = C.create_string_buffer(64)
        self._rc(self.lib.bh_ipc_export(self.h, mine))
        handles = [None] * world_size
        dist.all_gather_object(handles, mine.raw, group=group)
        blob = C.create_string_buffer(b"".join(handles), 64 * world_size)
        self._rc(self.lib.bh_ipc_set_peers(self.h, world_size, rank, blob))

    def disconnect_peers(self):
        self._rc(self.lib.bh_ipc_clear_peers(self.h))

    def set_slice(self, first, count):
        self._rc(self.lib.bh_set_slice(self.h, first, count))

    def step_fused_async(self, nsteps):
        """The whole sliced step inside the library (peer stores + device barrier, one CUDA graph per step)."""
        self._rc(self.lib.bh_step_async(self.h, nsteps))

    def force_slice(self, first, count):
        self._rc(self.lib.bh_calculate_force_slice(self.h, first, count))

    def apply_and_integrate(self):
        self._rc(self.lib.bh_finish_async(self.h))

    def check(self):
        self._rc(self.lib.bh_check(self.h))


class DistributedBarnesHutSimulation:
    """step() for world_size ranks; with world_size == 1 it degenerates to slice = everything."""

    def __init__(self, engine, rank: int, world_size: int, group=None):
        self.engine, self.rank, self.world_size, self.group = engine, rank, world_size, group
        self.chunk, self.bounds = slice_bounds(engine.nbodies, world_size)
        if self.chunk * world_size > engine.acc_sorted.shape[0]:
            raise ValueError("acceleration buffer too small for %d ranks" % world_size)
        self.fused = bool(getattr(engine, "p2p", False)) and world_size > 1
        self.peer_error = None
        if self.fused:
            # every rank must take the same path: agree on whether peer mapping worked everywhere
            import torch
            import torch.distributed as dist
            try:
                engine.connect_peers(rank, world_size, group)
                ok = 1
            except Exception as exc:  # no peer access between some pair of devices: use the NCCL all-gather
                self.peer_error = str(exc)
                ok = 0
            flag = torch.tensor([ok], device=engine.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            self.fused = bool(flag.item())
            engine.p2p = self.fused
            if self.fused:
                engine.set_slice(*self.bounds[rank])
            else:
                # ranks that did map their peers must unmap them: all ranks run the same (NCCL) protocol
                engine.disconnect_peers()

    def step_async(self, nsteps: int = 1):
        if self.fused:
            self.engine.step_fused_async(nsteps)
            return
        first, count = self.bounds[self.rank]
        full = self.engine.acc_sorted[: self.chunk * self.world_size]
        mine = full[self.rank * self.chunk:(self.rank + 1) * self.chunk]
        for _ in range(nsteps):
            self.engine.tree_stages()
            self.engine.force_slice(first, count)
            if self.world_size > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(full, mine, group=self.group)
            self.engine.apply_and_integrate()

    def step(self, nsteps: int = 1):
        self.step_async(nsteps)
        self.engine.check()
