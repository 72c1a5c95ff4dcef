#!/bin/bash
# 1-GPU box: GPU suite with the pruned z range on (default), C3 / C2 timings with and without it
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_prune_suite.log 2>&1; tail -4 gpurun_out/r2_prune_suite.log | cut -c1-400
for v in 0 1; do
  P3M_TUNE_NO_PRUNE=$v python bench.py --config mesh --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mesh no_prune=$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['ms_per_step_by_phase'].items() if x>0})"
done
python bench.py --steps 5 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-600
