set -u
for L in default lib_u4_w16 lib_u16_w16 lib_u8_w20 lib_u8_w24; do
  if [ $L = default ]; then unset P3M_B200_LIB; else export P3M_B200_LIB=$PWD/gpurun_alt/$L.so; fi
  timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ab2_$L.log 2>&1
  echo "$L: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab2_$L.log | head -1)"
done
