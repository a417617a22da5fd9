#!/bin/bash
# quick GPU check: parity tests + short bench (no CPU baseline)
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
echo "=== bench"; timeout 600 python bench.py --steps 50 --warmup 3 --skip-cpu --skip-65k --e2e-steps 1 $@ 2>&1 | tail -3 | tee gpurun_out/bench_quick.json
