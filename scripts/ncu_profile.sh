#!/bin/bash
# The ncu captures behind profiles/ (run on a B200 box; B200_PROFILING.md recipe).  Usage: scripts/ncu_profile.sh [outdir]
set -e
OUT=${1:-gpurun_out}; mkdir -p "$OUT"
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file "$OUT/launches_plummer10m.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
# full capture of the six step kernels of one timed step
ncu --set full --clock-control none --import-source on \
    -k regex:"force2_kernel|build_kernel|summarize_kernel|sort_kernel|integrate_kernel|bbox_kernel" -s 18 -c 6 \
    -o "$OUT/prof_plummer10m" python bench.py --steps 1 --warmup 3 --no-cpu > "$OUT/ncu_full.log" 2>&1
echo "read with: ncu -i $OUT/prof_plummer10m.ncu-rep --page raw --csv | --page source --csv"
