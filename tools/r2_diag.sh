#!/bin/bash
# 2-GPU box: multi-GPU parity tests, the coupled C5 sweep at N = 2 with per-rank phases, C2 A/B of the dense-cell threshold
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2_diag_smi.log 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/r2_diag_mgtests.log 2>&1; tail -5 gpurun_out/r2_diag_mgtests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r2_diag_g2.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2_diag_g2.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N=2 ms/step", d["ms_per_step"], "clocks", d["clocks"].get("per_rank"))
        for k, v in d["ms_per_step_by_phase_per_rank"].items(): print("  ", k, v)
        print("  parity", d["parity"]); print("  e2e", d["e2e"])
PY
tail -3 gpurun_out/r2_diag_g2.log | cut -c1-600
for dc in 64 32; do
  P3M_TUNE_DENSE_CELL=$dc timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_diag_c2_dense$dc.log 2>&1
  python - $dc <<'PY'
import json, sys
for l in open(f"gpurun_out/r2_diag_c2_dense{sys.argv[1]}.log"):
    if l.startswith("{"):
        d = json.loads(l); print("C2 dense", sys.argv[1], d["ms_per_step"], {k: round(v, 3) for k, v in d["ms_per_step_by_phase"].items()}, d["roofline"]["frac"], d["parity"]["sr_rel_l2"])
PY
done
