#!/bin/bash
mkdir -p gpurun_out
export DYNEMOL_B200_SERIES=blocked
timeout 400 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:blocked_series -s 2 -c 1 -f -o gpurun_out/prof_blocked \
    python bench.py --basis ${1:-4096} --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-65k > gpurun_out/ncu_blocked.log 2>&1
tail -3 gpurun_out/ncu_blocked.log
