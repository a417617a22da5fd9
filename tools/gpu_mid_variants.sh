#!/bin/bash
# phase profile + us per term of mid-kernel build variants (dynemol_b200/lib/libdyb_prof_*.so)
mkdir -p gpurun_out; : > gpurun_out/mid_variants.log
for lib in dynemol_b200/lib/libdyb_prof_*.so; do
  echo "=== $lib" | tee -a gpurun_out/mid_variants.log
  DYNEMOL_B200_LIB=$PWD/$lib timeout 200 python tools/gpu_mid.py --sizes ${SIZES:-2048,4096} --l2mb ${L2MB:-0} --reps 60 2>&1 | grep -E "mid_prof|per_term|rror" | tee -a gpurun_out/mid_variants.log
done
