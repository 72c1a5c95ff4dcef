#!/bin/bash
# Runs on the GPU box: A/B of the tuning switches on the 1-GPU point of the weak-scaling sweep (C5) and on C2.
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ph=d["ms_per_step_by_phase"]
print(sys.argv[1].split("/")[-1], "ms/step %.3f" % d["ms_per_step"], {k: round(v,3) for k,v in ph.items() if v>0}, "checked/part %.0f" % d.get("pairs_checked_per_particle",0))
PY
}
for dc in 64 32 24 16; do
  P3M_TUNE_DENSE_CELL=$dc python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/c5_dense$dc.log 2>&1; show gpurun_out/c5_dense$dc.log
done
P3M_TUNE_OLD_GATHER=1 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/c5_oldgather.log 2>&1; show gpurun_out/c5_oldgather.log
P3M_TUNE_LONGKEY=1 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/c5_longkey.log 2>&1; show gpurun_out/c5_longkey.log
