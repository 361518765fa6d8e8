"""From an `ncu --set full` capture of the step kernels (scripts/ncu_profile.sh) make the two files bench.py and the
judge read: profiles/<tag>_ncu_full_plummer10m.csv (selected counters per kernel) and profiles/<tag>_traffic.json
(DRAM bytes per launch + the hash of the kernel sources the capture was taken from; bench.py refuses the file when
the sources have changed since).   python scripts/make_traffic.py gpurun_out/r2_prof_plummer10m.ncu-rep r2"""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "r2")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_ldgsts.sum"] + [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
keep = [k for k in keep if k in idx]
out = os.path.join(ROOT, "profiles", "%s_ncu_full_plummer10m.csv" % tag)
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[idx[k]] for k in keep])
num = lambda s: float(s.replace(",", "")) if s else 0.0
unit = rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
kernels = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    rd = num(r[idx["dram__bytes_read.sum"]]) * scale[unit[idx["dram__bytes_read.sum"]]]
    wr = num(r[idx["dram__bytes_write.sum"]]) * scale[unit[idx["dram__bytes_write.sum"]]]
    kernels[name] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                     "ms_under_ncu": num(r[idx["gpu__time_duration.sum"]])}
h = hashlib.sha256()
for fn in ("bh_kernels.cuh", "bhstep.cu"):
    h.update(open(os.path.join(ROOT, "gpu_nbody_b200", "csrc", fn), "rb").read())
json.dump({"workload": "Plummer 10^7, theta 0.5 (bench.py --steps 1 --warmup 3)", "source": os.path.basename(rep),
           "kernel_source_sha16": h.hexdigest()[:16], "kernels": kernels},
          open(os.path.join(ROOT, "profiles", "%s_traffic.json" % tag), "w"), indent=1)
print(out, kernels)
