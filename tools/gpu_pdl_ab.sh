#!/bin/bash
mkdir -p gpurun_out
echo "=== tests (PDL on)"; timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for pdl in 0 1 0 1; do
  export DYNEMOL_B200_PDL=$pdl
  timeout 300 python bench.py --steps 60 --warmup 3 --skip-cpu --skip-e2e --skip-65k 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print('PDL',os.environ['DYNEMOL_B200_PDL'],'value',d['value'],'us/term',round(1e3*d['ms_per_step']/24,2),'kernel_us',r['kernel_avg_us'],'clk',d['clocks']['sm_mhz'])"
done
