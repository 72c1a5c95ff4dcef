#!/bin/bash
# 1-GPU box: r02 evidence -- launch lists of the bench command (C2) and of the mesh config (C3), one full ncu
# capture per hot kernel.  Summaries: python tools/summarize_ncu.py <tag>
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
C2="python bench.py --steps 2 --warmup 3 --quick"
C3="python bench.py --config mesh --steps 2 --warmup 3"
export P3M_TUNE_INC_SORT_DEN=3   # keep the mover-merge path on at the 11.5 % movers of this workload (default: full sort above n/12)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $C2 > gpurun_out/launches_${TAG}.stdout 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}mesh.csv $C3 > gpurun_out/launches_${TAG}mesh.stdout 2>&1
for K in k_pp_packed k_pp_sparse k_gather_p3m "k_deposit<"; do
  N=$(echo $K | tr -d '<')
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o /tmp/prof_${N}_${TAG} $C2 > gpurun_out/prof_${N}_${TAG}.stdout 2>&1
  ncu -i /tmp/prof_${N}_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${N}_${TAG}.raw.csv 2>/dev/null
done
# inner loop of the dominant kernel: source page (SASS with per-instruction counters), compressed
ncu -i /tmp/prof_k_pp_packed_${TAG}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/prof_k_pp_packed_${TAG}.source.csv.gz
for K in k_deposit_pm k_gather_pm k_poisson_z k_inc_place_stayers k_inc_place_movers k_rx_scatter; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o /tmp/prof_${K}_${TAG} $C3 > gpurun_out/prof_${K}_${TAG}.stdout 2>&1
  ncu -i /tmp/prof_${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${K}_${TAG}.raw.csv 2>/dev/null
done
du -sh gpurun_out
ls -la gpurun_out | grep ${TAG}
