set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short --timeout 600 -x > gpurun_out/pytest17.log 2>&1; grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest17.log | cut -c1-300 | tail -12
for SB in 3 4 5 6; do
  P3M_TUNE_SUBBITS=$SB timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_sub32_sb$SB.log 2>&1
  echo "SUB32 sbits=$SB: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_sub32_sb$SB.log | head -1) $(grep -o '"pairs_checked": [0-9]*' gpurun_out/ab_sub32_sb$SB.log)"
done
for SB in 3 5; do
  P3M_B200_LIB=$PWD/gpurun_alt/libp3m_b200_sub64.so P3M_TUNE_SUBBITS=$SB timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ab_sub64_sb$SB.log 2>&1
  echo "SUB64 sbits=$SB: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_sub64_sb$SB.log | head -1) $(grep -o '"pairs_checked": [0-9]*' gpurun_out/ab_sub64_sb$SB.log)"
done
