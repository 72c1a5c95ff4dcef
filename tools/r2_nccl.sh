#!/bin/bash
# 2-GPU box: effect of NCCL point-to-point channel counts on the exchanges of the C5 / C3 steps
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
run() { # name, config
  timeout 300 $TR bench.py --gpus 2 --config $2 --steps 5 --warmup 3 --no-extra > gpurun_out/r2_nccl_$1_$2.log 2>&1
  python - "$1" "$2" <<'PY'
import json, sys
for l in open(f"gpurun_out/r2_nccl_{sys.argv[1]}_{sys.argv[2]}.log"):
    if l.startswith("{"):
        d = json.loads(l); ex = d["exchange"]
        print(sys.argv[1], sys.argv[2], "ms/step %.3f" % d["ms_per_step"], "comm", ex["comm_ms_min_median_max"], "GB/s", round(ex["nvlink_GBs_per_rank_achieved"] or 0, 1), "MB", round(ex["bytes_sent_per_rank_per_step_max"] / 1e6, 1))
PY
}
run default c2; run default mesh
export NCCL_MIN_P2P_NCHANNELS=8; run p2p8 c2; run p2p8 mesh
export NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32; run p2p16 c2; run p2p16 mesh
export NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_MIN_NCHANNELS=32; run p2p32 mesh
