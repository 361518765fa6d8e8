#!/usr/bin/env python
"""bench.py -- body-steps/s of the Barnes-Hut step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--bodies B] [--dist plummer|uniform|disks]

One "step" = one full Barnes-Hut step (bounding box, tree build, summarise, sort,
force walk, integrate) over all bodies.  Default workload: BASELINE.json
configs[2] -- seeded Plummer sphere, 10 000 000 bodies, theta = 0.5 -- at every N
(strong scaling 1/2/4/8), which fits one GPU; `--bodies 1048576` gives configs[1].

`value`      whole-job body-steps/s, state resident in HBM (CUDA events, max over ranks).
`e2e`        the same step driven through the C ABI with HOST buffers: bh_upload (pinned
             host -> device) + bh_step + bh_copy_vertices (device -> host) every step.
`roofline`   the force kernel against the FP32 CUDA-core peak (flops = 20 I + 10 O,
             I/O counted by the instrumented kernel), plus the HBM-bound stages.
`cpu_baseline` the CPU oracle (port of the reference kernels) on this box's host cores.
--impl reference runs only that CPU arm (the reference's OpenCL/Java cannot run here:
no OpenCL CPU device, no JVM -- see DESIGN.md) and prints the same line shape.
"""
import argparse
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


def make_universe(dist, n, seed):
    from gpu_nbody_b200 import universe as U
    gen = {"plummer": lambda: U.PlummerUniverseGenerator(seed),
           "uniform": lambda: U.RandomCubicUniverseGenerator(6.0, seed),
           "disks": lambda: U.TwoDiskGalaxiesGenerator(seed, seed + 1)}[dist]()
    return U.generate_arrays(gen, n)


def workload_name(dist, n):
    return {"plummer": "Plummer sphere", "uniform": "uniform random cube (range 6)", "disks": "two colliding disk galaxies"}[dist] + \
        " %d bodies fp32, theta=0.5, eps2=0.0025, dt=0.025, vote width 16" % n


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


def cpu_oracle_sample(arrays, n, budget_s=20.0, steps=1, warmup=0):
    """The oracle's step on a bounded sample: all tree stages on the full body set,
    the force walk on `sample` sorted bodies starting at rotating offsets, the
    integrate on all bodies; per-step time = tree + force * n/sample + integrate."""
    import oracle
    oracle.use_all_cores()
    orc = oracle.OracleSim(n, *arrays, theta=THETA, eps2=EPS2, dt=DT, vote_width=16, fma_policy=1)
    t = time.perf_counter(); orc.bounding_box(); orc.build_tree(); orc.summarize(); orc.sort(); t_tree = time.perf_counter() - t
    # calibrate the sample so that one force sample takes about budget_s / (steps + warmup)
    probe = min(n, 16 * 1024)
    t = time.perf_counter(); orc.calculate_force_range(0, probe); t_probe = time.perf_counter() - t
    per_step = max(0.5, budget_s / max(1, steps + warmup))
    sample = int(min(n, max(probe, probe * per_step / max(t_probe, 1e-6))))
    sample -= sample % 16
    sample = max(16, sample)
    times, inter = [], 0
    for i in range(warmup + steps):
        first = 0 if sample >= n else ((i * sample) % (n - sample)) // 16 * 16
        t = time.perf_counter(); orc.calculate_force_range(first, min(sample, n - first)); tf = time.perf_counter() - t
        if i >= warmup:
            times.append(tf); inter += orc.interactions
    t = time.perf_counter(); orc.integrate(); t_int = time.perf_counter() - t
    t_force = float(np.mean(times)) * n / sample
    step_s = t_tree + t_force + t_int
    return {"value": n / step_s, "step_s": step_s, "tree_s": t_tree, "force_s_extrapolated": t_force, "integrate_s": t_int,
            "sample_bodies": sample, "cores": oracle.num_threads(),
            "sample": "tree build/summarise/sort + integrate on all %d bodies (sequential, timed once), force walk on %d consecutive sorted "
                      "bodies per step (OpenMP, %d threads) scaled by n/sample" % (n, sample, oracle.num_threads())}


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.bodies
    arrays = make_universe(args.dist, n, args.seed)
    res = cpu_oracle_sample(arrays, n, budget_s=60.0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "body-steps/sec", "value": res["value"], "unit": "body-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["step_s"] * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded %s, seed %d)" % (args.dist, args.seed),
            "config": {"workload": workload_name(args.dist, n), "bodies": n,
                       "note": "CPU oracle = C port of the reference's six OpenCL kernels (the OpenCL/Java reference cannot run on this image)"},
            "cpu_baseline": {"value": res["value"], "unit": "body-steps/s", "cores": res["cores"], "kind": "port", "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from gpu_nbody_b200 import GPUBarnesHutNBodySimulation, Mode, _lib, universe as U
    from gpu_nbody_b200.distributed import CudaSliceEngine, DistributedBarnesHutSimulation

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, K, W = args.bodies, args.steps, max(3, args.warmup)
    if args.gpu_gen:
        # memory-sized configurations (BASELINE configs[3], 10^8 bodies): the universe is drawn on the device
        # (bh_generate_universe: Philox, same seed on every rank = same bytes), no host arrays at all
        if args.dist == "disks":
            raise SystemExit("--gpu-gen supports --dist uniform and plummer (bh_generate_universe)")
        arrays = None
        sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, None, theta=THETA, eps2=EPS2, dt=DT, vote_width=16, device=local_rank)
        sim.init(None)
        sim.generateOnDevice("cubic" if args.dist == "uniform" else "plummer", args.seed, 6.0)
    else:
        arrays = make_universe(args.dist, n, args.seed)
        sim = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n, U.ArrayUniverseGenerator(*arrays), theta=THETA, eps2=EPS2, dt=DT, vote_width=16,
                                          device=local_rank)
        sim.init(None)
    lib = sim._lib
    engine = CudaSliceEngine(sim, p2p=(world > 1 and not args.nccl_allgather))  # also moves the simulation onto torch's current stream
    dsim = DistributedBarnesHutSimulation(engine, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up; the last warm-up step is counted (I, O of the roofline) ----
    inter = opens = 0
    if world == 1:
        sim.step(W - 1)
        sim.setCounting(True); sim.step(1); st = sim.stats(); sim.setCounting(False)
        inter, opens = st["interactions"], st["opens"]
    else:
        dsim.step(W)

    # ---- timed region: state resident in HBM ----
    sim.setProfiling(world == 1); sim.resetStats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    if world == 1:
        sim._check(lib.bh_step_async(sim.handle, K))
    else:
        dsim.step_async(K)
    e1.record()
    barrier()
    engine.check()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    st = sim.stats()
    launches = int(sum(st["stage_launches"].values()))
    stage_ms = {k: (v / st["steps_timed"] if st["steps_timed"] else None) for k, v in st["stage_ms"].items()}

    # ---- e2e: host buffers in, host buffers out, every step ----
    if args.gpu_gen or args.no_e2e:
        if rank == 0:
            value = n * K / (ms_total * 1e-3)
            print(json.dumps({"metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                              "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic (%s drawn on the device by bh_generate_universe, Philox seed %d)" % (args.dist, args.seed),
                              "config": {"workload": workload_name(args.dist, n), "bodies": n, "parallelism": "replicated tree, %d sorted slice(s)" % world},
                              "clocks": clocks, "gpu_launches": launches, "e2e": None, "cells_used": st["cells_used"], "max_depth": st["max_depth"],
                              "stage_ms": stage_ms}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in arrays]
    pos4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    vel4 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    sim.setProfiling(False)
    Ke = max(1, min(K, 5))

    def e2e_step():
        sim._check(lib.bh_upload(sim.handle, *(p.data_ptr() for p in pinned)))
        if world == 1:
            sim.step(1)
        else:
            dsim.step(1)
        if rank == 0:
            sim._check(lib.bh_copy_vertices(sim.handle, pos4.data_ptr(), vel4.data_ptr()))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n * Ke / float(t_e2e.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = n * K / (ms_total * 1e-3)
    line = {"metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded %s, seed %d)" % (args.dist, args.seed),
            "config": {"workload": workload_name(args.dist, n), "bodies": n, "parallelism": "replicated tree, %d sorted slice(s)%s" % (world, "" if world == 1 else (", all-gather fused into the force kernel (peer stores over NVLink)" if dsim.fused else ", NCCL all-gather")),
                       "l2": "working set (%.1f GB of tree + body state) is larger than the 126 MB L2; no flush needed" % (
                           (16 * (sim.numberOfNodes + 1) + 32 * n + 168 * (sim.numberOfNodes - n + 1) + 20 * n) / 1e9)},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "body-steps/s", "steps": Ke,
                    "h2d_bytes_per_step": 28 * n, "d2h_bytes_per_step": 32 * n},
            "cells_used": st["cells_used"], "max_depth": st["max_depth"]}
    if world == 1 and n != (1 << 20) and not args.no_cpu:
        # BASELINE configs[1] beside the headline workload: Plummer 2^20 on the same GPU, same protocol
        n1 = 1 << 20
        a1 = make_universe("plummer", n1, 42)
        s1 = GPUBarnesHutNBodySimulation(Mode.DEFAULT, n1, U.ArrayUniverseGenerator(*a1), theta=THETA, eps2=EPS2, dt=DT, vote_width=16,
                                         device=local_rank)
        s1.init(None)
        s1.setStream(torch.cuda.current_stream(dev).cuda_stream)
        s1.step(W)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record(); s1._check(lib.bh_step_async(s1.handle, K)); f1.record(); torch.cuda.synchronize()
        s1._check(lib.bh_check(s1.handle))
        line["configs1_plummer_1m"] = {"workload": workload_name("plummer", n1), "value": n1 * K / (f0.elapsed_time(f1) * 1e-3),
                                       "unit": "body-steps/s", "ms_per_step": f0.elapsed_time(f1) / K}
        s1.close()
    if world == 1:
        peak = SMS * LANES * 2 * 1.965e9 / 1e12
        peak_src = "computed 148 SM x 128 lanes x 2 x 1.965 GHz (FP32 CUDA-core peak is not in MEASURED_PEAKS.json)"
        meas = __import__("ctypes").c_double()
        if lib.bh_measure_fp32_peak(local_rank, __import__("ctypes").byref(meas)) == 0:
            line["fp32_peak_measured_tflops"] = meas.value
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        flops = 20.0 * inter + 10.0 * opens
        f_ms = stage_ms["calculate_force"]
        achieved = flops / (f_ms * 1e-3) / 1e12
        C = st["cells_used"]
        alg = {"bounding_box": 12 * n, "build_tree": 16 * n + 52 * C, "summarize": 16 * n + 104 * C, "sort": 4 * n + 44 * C,
               "integrate": 60 * n}
        traffic = {}
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this very workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if n == 10_000_000 and args.dist == "plummer":
                traffic = {k.split("<")[0]: v["dram_bytes_per_launch"] for k, v in tj["kernels"].items()}
        except Exception:
            pass
        line["roofline"] = {"kernel": "force2_kernel<16,false,false>", "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak, "traffic": traffic.get("force2_kernel"), "peak_source": peak_src,
                            "flops_per_launch": flops, "interactions_per_body": inter / n, "opens_per_body": opens / n,
                            "ms_per_launch": f_ms, "share_of_step": f_ms / sum(stage_ms.values())}
        kname = {"bounding_box": "bbox_kernel", "build_tree": "build_kernel", "summarize": "summarize_kernel", "sort": "sort_kernel",
                 "integrate": "integrate_kernel"}
        line["stages"] = {k: {"ms": stage_ms[k], "bound": "hbm", "alg_bytes": alg[k], "achieved_gbs": alg[k] / (stage_ms[k] * 1e-3) / 1e9,
                              "frac": alg[k] / (stage_ms[k] * 1e-3) / 1e9 / hbm_peak, "traffic": traffic.get(kname[k])} for k in alg}
        line["stages"]["hbm_peak_gbs"] = hbm_peak
        line["stages"]["hbm_peak_source"] = hbm_src
        if not args.no_cpu:
            res = cpu_oracle_sample(arrays, n, budget_s=15.0, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": res["value"], "unit": "body-steps/s", "cores": res["cores"], "kind": "port", "sample": res["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
