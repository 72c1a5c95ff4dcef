#!/bin/bash
# Runs on the GPU box (under gpurun): one full ncu capture of ONE kernel of the bench command.
#   tools/profile_kernel.sh <kernel regex> <tag> [bench args...]
# The .ncu-rep lands in gpurun_out/prof_<kernel>_<tag>.ncu-rep; tools/summarize_ncu.py <tag> condenses it.
set -u
mkdir -p gpurun_out
K=$1; TAG=$2; shift 2
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f \
    -o gpurun_out/prof_${K}_${TAG} $BENCH > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
tail -3 gpurun_out/prof_${K}_${TAG}.stdout
