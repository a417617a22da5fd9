#!/bin/bash
# ncu --set full of one launch of the mid-size series kernel (24 terms) at N = ${NCU_N:-4096}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mid_series -s 4 -c 1 -f -o gpurun_out/mid_ncu \
    python tools/gpu_mid.py --sizes ${NCU_N:-4096} --l2mb 0 --reps 4 2>&1 | tail -5
ls -la gpurun_out/
