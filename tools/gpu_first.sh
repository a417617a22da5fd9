#!/bin/bash
# first GPU contact: smoke, parity tests, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|Thread|Core" >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --maxfail=10 -x 2>&1 | tail -40
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --e2e-steps 1 --cpu-budget 8 2>&1 | tail -5 | tee gpurun_out/bench_first.json
