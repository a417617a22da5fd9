#!/bin/bash
# correctness + 3 bench repetitions of the current build
echo "=== tests"; timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for i in 1 2 3; do
  timeout 300 python bench.py --steps 60 --warmup 3 --skip-cpu --skip-e2e --skip-65k 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',d['value'],'us/term',round(1e3*d['ms_per_step']/24,2),'kernel_us',r['kernel_avg_us'],'GB/s',r['achieved'],'clk',d['clocks']['sm_mhz'])"
done
