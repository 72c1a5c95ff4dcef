#!/bin/bash
# 4-GPU box: the multi-GPU parity tests (4-rank case included) + the driver's N = 4 command
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
bash tools/r2_g8.sh 4
