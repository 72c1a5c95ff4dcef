#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one full ncu capture per hot
# kernel.  Outputs land in gpurun_out/; summaries are copied into profiles/ by tools/summarize_ncu.py.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.stdout 2>&1
for K in k_pp_tiled k_pp_sparse k_deposit k_gather; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} $BENCH > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
done
ls -la gpurun_out
