#!/bin/bash
# 1-GPU box: incremental-sort parity test, whole GPU suite, C3 / C5 / C2 timings with and without it
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "incremental_sort or short_key or sort_order" > gpurun_out/r2_sort_test.log 2>&1; tail -15 gpurun_out/r2_sort_test.log | cut -c1-300
show() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); ph=d["ms_per_step_by_phase"]
        print(sys.argv[1].split("/")[-1], "ms/step %.3f" % d["ms_per_step"], {k: round(v,3) for k,v in ph.items() if v>0}, {k: d.get("stats",{}).get(k) for k in ("sort_movers","full_sorts","incremental_sorts")})
PY
}
for cfg in mesh c5; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r2_sort_$cfg.log 2>&1; show gpurun_out/r2_sort_$cfg.log
  P3M_TUNE_FULL_SORT=1 timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r2_sort_${cfg}_full.log 2>&1; show gpurun_out/r2_sort_${cfg}_full.log
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_sort_suite.log 2>&1; tail -4 gpurun_out/r2_sort_suite.log | cut -c1-300
