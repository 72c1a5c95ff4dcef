#!/bin/bash
# 8-GPU box: exactly the driver's N = 8 command (coupled C5 weak point + extra.c3 + extra.c4), with a summary
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
SECONDS=0
timeout 420 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_g${N}.log 2> gpurun_out/r2_g${N}.err
echo "rc=$? elapsed: ${SECONDS}s"
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_g${N}.err | grep -v "^\s" | tail -15 | cut -c1-400
python - $N <<'PY'
import json, sys
for l in open(f"gpurun_out/r2_g{sys.argv[1]}.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("N", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "value %.4g" % d["value"], "e2e", d["e2e"])
        for k, v in d["ms_per_step_by_phase_min_median_max_over_ranks"].items(): print("  ", k, v)
        print("  exchange", {k: v for k, v in d["exchange"].items() if k != "note"})
        print("  parity", {k: v for k, v in d["parity"].items() if k not in ("sr_check",)})
        for name, x in d["extra"].items():
            print(" extra", name, x.get("workload"), "ms/step", x.get("ms_per_step"), x.get("weak_scaling_efficiency_vs_this"))
            for k, v in (x.get("ms_per_step_by_phase_min_median_max_over_ranks") or {}).items(): print("    ", k, v)
            if x.get("exchange"): print("    exchange", {k: v for k, v in x["exchange"].items() if k != "note"})
            if x.get("parity"): print("    parity", {k: v for k, v in x["parity"].items() if k not in ("sr_check",)})
PY
