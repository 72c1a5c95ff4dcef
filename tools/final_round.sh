#!/bin/bash
# Round-end GPU batch: GPU tests, the bench lines (ours + reference arm), ncu launch list and captures.
set -u
mkdir -p gpurun_out
TAG=${1:-r01n}
timeout 400 python -m pytest tests -q -m gpu --tb=short --timeout 200 > gpurun_out/pytest_final.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_final.log | cut -c1-300 | tail -8
timeout 300 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-700
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; tail -1 gpurun_out/bench_ref_final.log | cut -c1-500
timeout 200 python bench.py --config mesh --steps 3 --warmup 2 > gpurun_out/bench_mesh_final.log 2>&1; tail -1 gpurun_out/bench_mesh_final.log | cut -c1-1500
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/launches_${TAG}.stdout 2>&1
for K in k_pp_tiled k_pp_sparse; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} $BENCH > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
done
MB="python bench.py --config mesh --steps 1 --warmup 1"
for K in k_gather_pm k_deposit_pm k_poisson_z; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f \
      -o gpurun_out/prof_${K}_${TAG} $MB > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
done
ls gpurun_out | grep ${TAG}
