#!/bin/bash
# mid-size series kernel: parity tests (bounded), phase profile (diagnostic build), us per term
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mid_kernel.py -x -q 2>&1 | tail -15 | tee gpurun_out/mid_tests.log
DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/libdyb_prof.so timeout 200 python tools/gpu_mid.py --sizes ${SIZES:-2048,4096} --l2mb ${L2MB:-0} --reps 60 2>&1 | grep -E "mid_prof|rror" | tee gpurun_out/mid_prof.log
timeout 300 python tools/gpu_mid.py --sizes ${SIZES2:-2048,2304,3000,4096,5000,6144} --l2mb ${L2MB2:-0} --reps 100 --out gpurun_out/mid_times.json 2>&1 | tail -8
if [ -f dynemol_b200/lib/libdyb_s0.so ]; then echo "--- DYB_LL_SLEEP=0"; DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/libdyb_s0.so timeout 300 python tools/gpu_mid.py --sizes ${SIZES2:-2048,2304,3000,4096,5000,6144} --l2mb ${L2MB2:-0} --reps 100 2>&1 | tail -8; fi
