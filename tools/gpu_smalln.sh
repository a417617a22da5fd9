#!/bin/bash
for mt in 1 2 4 8; do
  export DYNEMOL_B200_MIN_TILES=$mt
  for n in 512 1024 2048 4096; do
  timeout 200 python bench.py --basis $n --steps 200 --warmup 5 --skip-cpu --skip-65k --skip-e2e 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print('min_tiles',os.environ['DYNEMOL_B200_MIN_TILES'],'N',d['config']['basis'],'grid',d['config']['grid'],'us/term',round(1e3*d['ms_per_step']/24,2),'kernel_us',r['kernel_avg_us'])"
  done
done
