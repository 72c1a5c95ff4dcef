#!/bin/bash
# 1-GPU box: short-range parity tests + C5 / C2 timings over the dense-cell threshold
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_direct_sum.py -x -q -k "short_range or packed or force or direct or determin or unequal" 2>&1 | tail -4
for dc in 32 16 8; do
  P3M_TUNE_DENSE_CELL=$dc python bench.py --config c5 --steps 5 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 dense $dc', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['ms_per_step_by_phase'].items() if v>0}, round(d['pairs_checked_per_particle']), round(d['pairs_in_range_per_particle']))"
done
for dc in 32 8; do
  P3M_TUNE_DENSE_CELL=$dc python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2 dense $dc', round(d['ms_per_step'],3), round(d['ms_per_step_by_phase']['shortRangeForcesCalc'],3), d['roofline']['frac'], d['parity']['sr_rel_l2'])"
done
