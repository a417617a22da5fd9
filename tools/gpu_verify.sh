#!/bin/bash
# what the driver does at round end (no ncu): tests, smoke, reference arm, bench
R=${1:-r1f}
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_$R.json | cut -c1-200
echo "=== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_$R.json
