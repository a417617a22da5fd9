#!/bin/bash
# phase profile of the one-launch-per-series kernel (diagnostic build with -DDYB_SERIES_PROF)
mkdir -p gpurun_out
export DYNEMOL_B200_SERIES=stream DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/libdyb_prof.so
for n in 512 4096 16384; do
  timeout 200 python bench.py --basis $n --steps 40 --warmup 4 --skip-cpu --skip-65k --skip-e2e 2>&1 | grep -E "series_prof" | tail -3 | tee -a gpurun_out/persist_prof.log
done
