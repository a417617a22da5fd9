#!/bin/bash
# correctness + bench at several sizes for the current build
echo "=== tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for n in 16384 16384 512 2048 4096; do
  timeout 300 python bench.py --basis $n --steps 60 --warmup 3 --skip-cpu --skip-e2e --skip-65k 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N',d['config']['basis'],'value',d['value'],'us/term',round(1e3*d['ms_per_step']/24,2),'kernel_us',r['kernel_avg_us'],'GB/s',r['achieved'],'clk',d['clocks']['sm_mhz'])"
done
