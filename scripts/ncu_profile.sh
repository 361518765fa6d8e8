#!/bin/bash
# The ncu captures behind profiles/ (run on a B200 box; B200_PROFILING.md recipe).  Usage: scripts/ncu_profile.sh [outdir] [tag]
set -e
OUT=${1:-gpurun_out}; TAG=${2:-r2}; mkdir -p "$OUT"
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file "$OUT/${TAG}_launches_plummer10m.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > /dev/null 2>&1
# full capture of the six step kernels of the timed step (3 warm-up steps x 6 launches skipped)
ncu --set full --clock-control none --import-source on \
    -k regex:"walk_kernel|build_kernel|summarize_kernel|sort_kernel|finish_kernel|bbox_kernel" -s 18 -c 6 \
    -o "$OUT/${TAG}_prof_plummer10m" python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-extras > "$OUT/${TAG}_ncu_full.log" 2>&1
echo "read with: ncu -i $OUT/${TAG}_prof_plummer10m.ncu-rep --page raw --csv | --page source --csv; then scripts/make_traffic.py"
