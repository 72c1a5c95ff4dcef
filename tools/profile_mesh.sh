#!/bin/bash
# Runs on the GPU box: full ncu capture of the mesh-side kernels on the C3-style PM config.
set -u
mkdir -p gpurun_out
TAG=${1:-r01m}
BENCH="python bench.py --config mesh --steps 1 --warmup 1"
for K in k_gather_pm k_deposit k_poisson_z; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} $BENCH > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
done
