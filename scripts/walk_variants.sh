#!/bin/bash
# Build alternative tunings of the force walk (batch, sub-batch, refill trips, stack, CTAs/SM) into gpu_nbody_b200/variants/
# (run here, CPU box), then time them on the GPU box:
#   for f in gpu_nbody_b200/variants/*.so; do BHSTEP_LIBRARY=$f python bench.py --steps 5 --no-cpu --no-e2e --no-extras; done
set -e
cd "$(dirname "$0")/.."
mkdir -p gpu_nbody_b200/variants
build() {  # name batch sub trips scap ctas
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include \
       -DBH_WALK_BATCH=$2 -DBH_WALK_SUB=$3 -DBH_WALK_TRIPS=$4 -DBH_WALK_SCAP=$5 -DBH_WALK_CTAS=$6 $7 -Xptxas -v \
       -o gpu_nbody_b200/variants/$1.so gpu_nbody_b200/csrc/bhstep.cu 2>&1 | grep -A2 "walk_kernelILb0" | grep Used | sed "s/^/$1: /"
}
build b16s8t5c4 16 8 5 96 4 &
build b12s6t5c4 12 6 5 96 4 &
build b20s10t7c4 20 10 7 96 4 &
build b16s8t5c5 16 8 5 96 5 &
wait
