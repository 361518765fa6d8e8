#!/usr/bin/env python
"""bench.py -- body-steps/s of the Barnes-Hut step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--bodies B] [--dist plummer|uniform|disks]

One "step" = one full Barnes-Hut step (bounding box, tree build, summarise, sort,
force walk, integrate) over all bodies.  Default workload: BASELINE.json
configs[2] -- seeded Plummer sphere, 10 000 000 bodies, theta = 0.5 -- at every N
(strong scaling 1/2/4/8), which fits one GPU; `--bodies 1048576` gives configs[1].

`value`      whole-job body-steps/s, state resident in HBM (CUDA events, max over ranks).
`e2e`        the same step driven through the C ABI with HOST buffers: bh_upload_async (pinned
             host -> device) + bh_step_async + bh_copy_vertices_async (device -> host) every
             step, wall clock, the last read-back awaited inside the timed region.
`gpu_launches` stage kernels launched in the timed region (six per step and rank, one more with
             the peer barrier; the one-thread ticket reset before each walk is not counted).
`roofline`   the force kernel against the FP32 CUDA-core peak (flops = 20 I + 10 O,
             I/O counted by the instrumented kernel), plus the HBM-bound stages.
`cpu_baseline` the CPU oracle (port of the reference kernels) on this box's host cores:
             one complete step over all bodies.
N > 1 adds `stage_ms` (CUDA events around every stage and the peer barrier, taken in a
separate short pass: the timed region replays one CUDA graph per step) and
`parity_check`: after the timed steps rank 0 runs the same number of steps on ONE GPU
from the same input and compares positions, velocities, accelerations and the sorted
order bitwise with the distributed state; all ranks compare a hash of theirs.  A
mismatch is a non-zero exit.
--impl reference runs only the CPU arm (the reference's OpenCL/Java cannot run here:
no OpenCL CPU device, no JVM -- see DESIGN.md): complete oracle steps over all bodies,
as many as fit its time budget (`steps` = the number actually timed).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

THETA, EPS2, DT = 0.5, 0.0025, 0.025
SMS, LANES = 148, 128
STATE_KEYS = ("posX", "posY", "posZ", "velX", "velY", "velZ", "accX", "accY", "accZ", "sorted")
KERNEL_SOURCES = [os.path.join(ROOT, "gpu_nbody_b200", "csrc", f) for f in ("bh_kernels.cuh", "bhstep.cu")]


def make_universe(dist, n, seed):
    from gpu_nbody_b200 import universe as U
    gen = {"plummer": lambda: U.PlummerUniverseGenerator(seed),
           "uniform": lambda: U.RandomCubicUniverseGenerator(6.0, seed),
           "disks": lambda: U.TwoDiskGalaxiesGenerator(seed, seed + 1)}[dist]()
    return U.generate_arrays(gen, n)


def workload_name(dist, n, theta=THETA):
    return {"plummer": "Plummer sphere", "uniform": "uniform random cube (range 6)", "disks": "two colliding disk galaxies"}[dist] + \
        " %d bodies fp32, theta=%g, eps2=0.0025, dt=0.025, vote width 16" % (n, theta)


def kernel_source_hash():
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        # samples taken while the GPU was busy: the upper half of the clock samples
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C port of the reference's six kernels), complete steps over all bodies
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_steps(arrays, n, max_steps, warmup, budget_s):
    """Complete oracle steps over all n bodies: `warmup` untimed, then up to `max_steps` timed ones, stopping when
    the next step would exceed `budget_s` (at least one is always timed).  Nothing is sampled or extrapolated."""
    import oracle
    oracle.use_all_cores()
    orc = oracle.OracleSim(n, *arrays, theta=THETA, eps2=EPS2, dt=DT, vote_width=16, fma_policy=1)
    stage_fns = (("bounding_box", orc.bounding_box), ("build_tree", orc.build_tree_parallel), ("summarize", orc.summarize),
                 ("sort", orc.sort), ("calculate_force", orc.calculate_force), ("integrate", orc.integrate))
    t_start = time.perf_counter()
    for _ in range(warmup):
        for _name, fn in stage_fns:
            fn()
    times, stages = [], None
    while len(times) < max_steps:
        t0 = time.perf_counter()
        st = {}
        for name, fn in stage_fns:
            t = time.perf_counter(); fn(); st[name] = time.perf_counter() - t
        times.append(time.perf_counter() - t0)
        stages = st if stages is None else {k: stages[k] + st[k] for k in st}
        if time.perf_counter() - t_start + float(np.mean(times)) > budget_s:
            break
    step_s = float(np.mean(times))
    return {"value": n / step_s, "step_s": step_s, "steps": len(times), "warmup": warmup, "cores": oracle.num_threads(),
            "stage_s": {k: v / len(times) for k, v in stages.items()}, "measured_s": float(np.sum(times)),
            "sample": "%d complete step(s) over all %d bodies, timed whole (nothing extrapolated); bounding box, tree build (concurrent "
                      "insertion with CAS locks, as in buildtree.cl), force walk and integrate on %d OpenMP threads; summarise and sort "
                      "(4 %% of the step) sequential" % (len(times), n, oracle.num_threads())}


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.bodies
    arrays = make_universe(args.dist, n, args.seed)
    res = cpu_oracle_steps(arrays, n, max_steps=args.steps, warmup=min(args.warmup, 1), budget_s=args.cpu_budget)
    line = {"impl": "reference", "metric": "body-steps/sec", "value": res["value"], "unit": "body-steps/s", "n_gpus": args.gpus,
            "steps": res["steps"], "warmup": res["warmup"], "requested_steps": args.steps, "requested_warmup": args.warmup,
            "extrapolated": False, "measured_s": res["measured_s"],
            "ms_per_step": res["step_s"] * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded %s, seed %d)" % (args.dist, args.seed),
            "config": {"workload": workload_name(args.dist, n), "bodies": n,
                       "note": "CPU oracle = C port of the reference's six OpenCL kernels (the OpenCL/Java reference cannot run on this image); "
                               "complete steps over all bodies, as many as fit %g s" % args.cpu_budget},
            "stage_ms": {k: v * 1e3 for k, v in res["stage_s"].items()},
            "cpu_baseline": {"value": res["value"], "unit": "body-steps/s", "cores": res["cores"], "kind": "port", "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
class Runner:
    """One simulation (single- or multi-GPU) with the bench protocol: warm-up, timed steps, stage pass, parity check."""

    def __init__(self, args, rank, world, local_rank, n, dist_name, seed, theta=THETA, gpu_gen=False, arrays=None):
        import torch
        from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
        from gpu_nbody_b200.distributed import CudaSliceEngine, DistributedBarnesHutSimulation
        self.torch, self.args, self.rank, self.world, self.local_rank = torch, args, rank, world, local_rank
        self.n, self.dist_name, self.seed, self.theta, self.gpu_gen = n, dist_name, seed, theta, gpu_gen
        self.dev = torch.device("cuda", local_rank)
        self.arrays = arrays
        if not gpu_gen and arrays is None:
            self.arrays = make_universe(dist_name, n, seed)
        self.sim = self._make_sim(local_rank)
        self.lib = self.sim._lib
        self.engine = CudaSliceEngine(self.sim, p2p=(world > 1 and not args.nccl_allgather))  # also moves the simulation onto torch's current stream
        self.dsim = DistributedBarnesHutSimulation(self.engine, rank, world)
        self.steps_done = 0

    def _make_sim(self, device):
        from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, universe as U
        gen = None if self.gpu_gen else U.ArrayUniverseGenerator(*self.arrays)
        sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, self.n, gen, theta=self.theta, eps2=EPS2, dt=DT, vote_width=16, device=device)
        sim.init(None)
        if self.gpu_gen:
            # memory-sized configurations (BASELINE configs[3], 10^8 bodies): the universe is drawn on the device
            # (bh_generate_universe: Philox, same seed on every rank = same bytes), no host arrays at all
            sim.generateOnDevice("cubic" if self.dist_name == "uniform" else "plummer", self.seed, 6.0)
        return sim

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def step(self, k):
        if self.world == 1:
            self.sim.step(k)
        else:
            self.dsim.step(k)
        self.steps_done += k

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.torch.distributed.all_reduce(t, op=self.torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def warmup(self, W, count=False):
        """W untimed steps; with count=True the last one runs the counting variant of the walk (I, O of the roofline)."""
        inter = opens = 0
        if count and self.world == 1:
            self.step(W - 1)
            self.sim.setCounting(True); self.step(1); st = self.sim.stats(); self.sim.setCounting(False)
            inter, opens = st["interactions"], st["opens"]
        else:
            self.step(W)
        return inter, opens

    def timed(self, K, profile=False, sampler=None):
        """K steps between two CUDA events on the stream the kernels run on; max over ranks."""
        torch = self.torch
        self.sim.setProfiling(profile); self.sim.resetStats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        self.barrier()
        e0.record()
        if self.world == 1:
            self.sim._check(self.lib.bh_step_async(self.sim.handle, K))
        else:
            self.dsim.step_async(K)
        e1.record()
        self.barrier()
        self.engine.check()
        self.steps_done += K
        clocks = sampler.stop() if sampler else None
        ms_total = self.max_over_ranks(e0.elapsed_time(e1))
        st = self.sim.stats()
        self.sim.setProfiling(False)
        return ms_total, st, clocks

    def stage_pass(self, P=3):
        """Per-stage CUDA-event times (and the peer barrier's) over P steps launched kernel by kernel."""
        self.sim.setProfiling(True); self.sim.resetStats()
        self.barrier()
        self.step(P)
        st = self.sim.stats()
        self.sim.setProfiling(False)
        ms = {k: v / max(1, st["steps_timed"]) for k, v in st["stage_ms"].items()}
        ms["peer_barrier"] = st["barrier_ms"] / max(1, st["steps_timed"])
        return {k: self.max_over_ranks(v) for k, v in ms.items()}  # same keys in the same order on every rank

    def state_hash(self, keys=STATE_KEYS):
        h = hashlib.sha256()
        for k in keys:
            h.update(self.sim.readBuffer(k, self.n).tobytes())
        return h.digest()

    def parity_check(self, keys=STATE_KEYS):
        """Distributed state == single-GPU state after the same number of steps (bitwise), and identical on all ranks."""
        torch = self.torch
        mine = self.state_hash(keys)
        word = torch.tensor([int.from_bytes(mine[:8], "little", signed=True)], device=self.dev, dtype=torch.int64)
        words = [torch.zeros_like(word) for _ in range(self.world)]
        torch.distributed.all_gather(words, word)
        identical = all(int(w.item()) == int(word.item()) for w in words)
        ok = None
        if self.rank == 0:
            ref = self._make_sim(self.local_rank)
            ref.step(self.steps_done)
            h = hashlib.sha256()
            for k in keys:
                h.update(ref.readBuffer(k, self.n).tobytes())
            ok = h.digest() == mine
            ref.close()
        flag = torch.tensor([1 if (ok is None or ok) else 0], device=self.dev, dtype=torch.int32)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        return {"vs_single_gpu_bitwise": bool(flag.item()), "ranks_identical": identical, "steps_compared": self.steps_done,
                "buffers": list(keys)}

    def close(self):
        self.sim.close()


def extra_config(args, rank, world, local_rank, key, n, dist_name, seed, theta, gpu_gen, K, W):
    """A BASELINE config beside the headline workload, same protocol (state resident, CUDA events, max over ranks)."""
    r = Runner(args, rank, world, local_rank, n, dist_name, seed, theta=theta, gpu_gen=gpu_gen)
    r.warmup(W)
    ms_total, st, clocks = r.timed(K, sampler=ClockSampler(local_rank) if rank == 0 else None)
    out = {"workload": workload_name(dist_name, n, theta), "value": n * K / (ms_total * 1e-3), "unit": "body-steps/s",
           "ms_per_step": ms_total / K, "n_gpus": world, "steps": K, "warmup": W, "cells_used": st["cells_used"],
           "max_depth": st["max_depth"], "clocks": clocks,
           "data": "synthetic (%s)" % ("drawn on the device, Philox seed %d" % seed if gpu_gen else "seeded, seed %d" % seed)}
    if world > 1:
        out["parity_check"] = r.parity_check(("posX", "velX", "accX", "sorted") if n > 20_000_000 else STATE_KEYS)
    r.close()
    return out


def run_ours(args, rank, world, local_rank):
    import ctypes as C
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, K, W = args.bodies, args.steps, max(3, args.warmup)
    if args.gpu_gen and args.dist == "disks":
        raise SystemExit("--gpu-gen supports --dist uniform and plummer (bh_generate_universe)")
    run = Runner(args, rank, world, local_rank, n, args.dist, args.seed, gpu_gen=args.gpu_gen)
    sim, lib = run.sim, run.lib

    # ---- warm-up; at N = 1 the last warm-up step is counted (I, O of the roofline) ----
    inter, opens = run.warmup(W, count=True)
    # ---- timed region: state resident in HBM.  N = 1: stage events on (kernel-by-kernel launches); N > 1: graph replay ----
    ms_total, st, clocks = run.timed(K, profile=(world == 1), sampler=ClockSampler(local_rank) if rank == 0 else None)
    launches = int(sum(st["stage_launches"].values()))
    if world == 1:
        stage_ms = {k: (v / st["steps_timed"] if st["steps_timed"] else None) for k, v in st["stage_ms"].items()}
    else:
        stage_ms = run.stage_pass(3)
    value = n * K / (ms_total * 1e-3)
    line = {"metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": ("synthetic (%s drawn on the device by bh_generate_universe, Philox seed %d)" if args.gpu_gen else "synthetic (seeded %s, seed %d)") % (args.dist, args.seed),
            "config": {"workload": workload_name(args.dist, n), "bodies": n,
                       "parallelism": "replicated tree, %d sorted slice(s)%s" % (world, "" if world == 1 else (
                           ", all-gather fused into the force kernel (peer stores over NVLink), device-side peer barrier, one CUDA graph per step"
                           if run.dsim.fused else ", NCCL all-gather")),
                       "l2": "working set (%.1f GB of tree + body state) is larger than the 126 MB L2; no flush needed" % (
                           (96 * n + 300 * (sim.numberOfNodes - n + 1)) / 1e9)},
            "clocks": clocks, "gpu_launches": launches, "stage_ms": stage_ms, "cells_used": st["cells_used"], "max_depth": st["max_depth"]}

    # ---- N > 1: the distributed state against a single-GPU run of the same steps ----
    parity_ok = True
    if world > 1:
        pc = run.parity_check()
        line["parity_check"] = pc
        parity_ok = pc["vs_single_gpu_bitwise"] and pc["ranks_identical"]

    # ---- e2e: host buffers in, host buffers out, every step ----
    if not (args.gpu_gen or args.no_e2e):
        arrays = run.arrays
        pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in arrays]
        pos4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        vel4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()

        def e2e_step():
            # pinned host arrays in (bh_upload_async: all copies on a second stream; the step starts when positions and
            # masses have arrived, its last pass when the velocities have), one step (bh_step_async), float4 vertices out to
            # pinned host buffers (bh_copy_vertices_async: read-back behind the next upload's position copies).  Nothing
            # waits between steps, so the next step's inputs cross the link while the current step computes.
            sim._check(lib.bh_upload_async(sim.handle, *(p.data_ptr() for p in pinned)))
            if world == 1:
                sim._check(lib.bh_step_async(sim.handle, 1))
            else:
                run.dsim.step_async(1)
            run.steps_done += 1
            if rank == 0:
                sim._check(lib.bh_copy_vertices_async(sim.handle, pos4.data_ptr(), vel4.data_ptr()))
        e2e_step()
        sim._check(lib.bh_wait_copies(sim.handle))
        run.engine.check()
        run.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            e2e_step()
        sim._check(lib.bh_wait_copies(sim.handle))  # the last step's vertices have arrived
        run.engine.check()                          # ... and every step ran without a device error
        run.barrier()
        t_e2e = run.max_over_ranks(time.perf_counter() - t0)
        line["e2e"] = {"value": n * K / t_e2e, "unit": "body-steps/s", "steps": K, "ms_per_step": 1e3 * t_e2e / K,
                       "h2d_bytes_per_step": 28 * n, "d2h_bytes_per_step": 32 * n}
    else:
        line["e2e"] = None

    # ---- the other BASELINE configs beside the headline (outside its timing; each with its own clock record) ----
    if args.extras and not args.gpu_gen and args.dist == "plummer" and n == 10_000_000:
        Ke, We = max(3, min(K, 10)), 3
        if world == 1:
            line["configs1_plummer_1m"] = extra_config(args, rank, world, local_rank, "configs1", 1 << 20, "plummer", 42, THETA, False, Ke, We)
            line["configs3_uniform_10m"] = extra_config(args, rank, world, local_rank, "configs3", 10_000_000, "uniform", 44, THETA, True, Ke, We)
            for th in (0.3, 0.5, 0.8):
                line["configs4_disks_4m_theta%02d" % round(th * 10)] = extra_config(args, rank, world, local_rank, "configs4", 4_000_000,
                                                                                     "disks", 45, th, False, Ke, We)
        else:
            line["configs3_uniform_10m"] = extra_config(args, rank, world, local_rank, "configs3", 10_000_000, "uniform", 44, THETA, True, Ke, We)
            if world == 8:
                line["configs3_uniform_100m"] = extra_config(args, rank, world, local_rank, "configs3", 100_000_000, "uniform", 44, THETA, True,
                                                             5, We)

    if rank != 0:
        run.close()
        if world > 1:
            dist.destroy_process_group()
        if not parity_ok:
            raise SystemExit(3)
        return

    if world == 1:
        computed = SMS * LANES * 2 * 1.965e9 / 1e12
        meas = C.c_double()
        peak, peak_src = computed, "computed 148 SM x 128 lanes x 2 x 1.965 GHz (FP32 CUDA-core peak is not in MEASURED_PEAKS.json)"
        if lib.bh_measure_fp32_peak(local_rank, C.byref(meas)) == 0 and meas.value > 0:
            peak, peak_src = meas.value, "measured FFMA microbenchmark, this run (bh_measure_fp32_peak; MEASURED_PEAKS.json has no FP32 CUDA-core figure)"
        x2 = C.c_double()
        x2_rate = x2.value if lib.bh_measure_fp32x2_rate(local_rank, C.byref(x2)) == 0 else None
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        flops = 20.0 * inter + 10.0 * opens
        f_ms = stage_ms["calculate_force"]
        achieved = flops / (f_ms * 1e-3) / 1e12
        Cc = st["cells_used"]
        alg = {"bounding_box": 12 * n, "build_tree": 16 * n + 52 * Cc, "summarize": 16 * n + 104 * Cc, "sort": 4 * n + 44 * Cc,
               "integrate": 60 * n}
        traffic, traffic_note = {}, None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this very workload and these very kernels
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if n == 10_000_000 and args.dist == "plummer":
                if tj.get("kernel_source_sha16") == kernel_source_hash():
                    traffic = {k: v["dram_bytes_per_launch"] for k, v in tj["kernels"].items()}
                else:
                    traffic_note = "profiles/r2_traffic.json was captured from other kernel sources (sha %s, now %s): not used" % (
                        tj.get("kernel_source_sha16"), kernel_source_hash())
        except Exception:
            traffic_note = "no profiles/r2_traffic.json"
        line["roofline"] = {"kernel": "walk_kernel<false>", "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak, "traffic": traffic.get("walk_kernel"), "peak_source": peak_src,
                            "peak_computed_tflops": computed, "frac_of_computed_peak": achieved / computed,
                            "ffma2_three_register_operands_measured_tflops": x2_rate,
                            "flops_per_launch": flops, "interactions_per_body": inter / n, "opens_per_body": opens / n,
                            "ms_per_launch": f_ms, "share_of_step": f_ms / sum(stage_ms.values()),
                            "instruction_mix_ceiling": "13 fp32-pipe lane-ops + 1 MUFU per 20 counted flops: at most 0.77 of the FMA peak",
                            # from the committed ncu capture and microbenchmarks (profiles/r2_microbench_fp32x2.txt), not measured in this run
                            "pipe_cycles_per_child_test": {"fma": 52, "xu": 32, "alu": 28, "issue_slots": 54, "elapsed": 82.6,
                                                           "note": "per SM sub-partition; no pipe saturated, four in-order warps per sub-partition"}}
        if traffic_note:
            line["roofline"]["traffic_note"] = traffic_note
        kname = {"bounding_box": "bbox_kernel", "build_tree": "build_kernel", "summarize": "summarize_kernel", "sort": "sort_kernel",
                 "integrate": "finish_kernel"}
        line["stages"] = {k: {"ms": stage_ms[k], "bound": "hbm", "alg_bytes": alg[k], "achieved_gbs": alg[k] / (stage_ms[k] * 1e-3) / 1e9,
                              "frac": alg[k] / (stage_ms[k] * 1e-3) / 1e9 / hbm_peak, "traffic": traffic.get(kname[k])} for k in alg}
        line["stages"]["hbm_peak_gbs"] = hbm_peak
        line["stages"]["hbm_peak_source"] = hbm_src
        if not args.no_cpu and not args.gpu_gen:
            res = cpu_oracle_steps(run.arrays, n, max_steps=1, warmup=0, budget_s=0.0)
            line["cpu_baseline"] = {"value": res["value"], "unit": "body-steps/s", "cores": res["cores"], "kind": "port", "sample": res["sample"],
                                    "stage_ms": {k: v * 1e3 for k, v in res["stage_s"].items()}}
    run.close()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", type=int, default=10_000_000)
    ap.add_argument("--dist", default="plummer", choices=["plummer", "uniform", "disks"])
    ap.add_argument("--seed", type=int, default=43)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (profiling runs)")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the other BASELINE configs beside the headline workload")
    ap.add_argument("--cpu-budget", type=float, default=100.0, help="--impl reference: seconds of complete oracle steps to time")
    ap.add_argument("--gpu-gen", action="store_true", help="draw the universe on the device (memory-sized runs; skips the e2e and cpu legs)")
    ap.add_argument("--nccl-allgather", action="store_true", help="multi-GPU: NCCL all-gather instead of the peer-memory stores fused into the force kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-launch ourselves one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
