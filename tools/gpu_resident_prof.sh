#!/bin/bash
# phase profile of the shared-memory-resident series kernel (diagnostic build with -DDYB_SERIES_PROF)
mkdir -p gpurun_out
export DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/libdyb_prof.so
for n in ${SIZES:-128 512 900 1792}; do
  timeout 200 python bench.py --basis $n --steps 60 --warmup 4 --skip-cpu --skip-65k --skip-e2e 2>&1 | grep -E "resident_prof" | tail -2 | tee -a gpurun_out/resident_prof.log
done
