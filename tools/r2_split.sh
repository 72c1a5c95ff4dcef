#!/bin/bash
# 2-GPU box: multi-GPU parity worker + exchanges with split / contiguous FFT slabs
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
run() { # name, config
  timeout 300 $TR bench.py --gpus 2 --config $2 --steps 5 --warmup 3 --no-extra > gpurun_out/r2_split_$1_$2.log 2>&1
  python - "$1" "$2" <<'PY'
import json, sys
for l in open(f"gpurun_out/r2_split_{sys.argv[1]}_{sys.argv[2]}.log"):
    if l.startswith("{"):
        d = json.loads(l); ex = d["exchange"]
        print(sys.argv[1], sys.argv[2], "ms/step %.3f" % d["ms_per_step"], "comm", ex["comm_ms_min_median_max"], "MB", round(ex["bytes_sent_per_rank_per_step_max"] / 1e6, 1), "parity", {k: v for k, v in d["parity"].items() if "rel" in k or "ident" in k})
PY
  grep -E "Error|error" gpurun_out/r2_split_$1_$2.log | head -3
}
run split c2; run split mesh
export P3M_TUNE_CONTIG_SLABS=1; run contig c2; run contig mesh
