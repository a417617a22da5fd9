#!/bin/bash
# phase profile of the mid-size series kernel (diagnostic build with -DDYB_SERIES_PROF) + timing of the shipped build
mkdir -p gpurun_out
DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/libdyb_prof.so timeout 300 python tools/gpu_mid.py --sizes ${SIZES:-2048,3000,4096,6144} --l2mb ${L2MB:-96} --reps 60 2>&1 | grep -E "mid_prof|per_term" | tee gpurun_out/mid_prof.log
timeout 300 python tools/gpu_mid.py --sizes ${SIZES:-2048,3000,4096,6144} --l2mb ${L2MB2:-0,64,96,120} --reps 100 --out gpurun_out/mid_times.json 2>&1 | tail -8
