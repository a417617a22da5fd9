#!/bin/bash
mkdir -p gpurun_out
for n in 128 512 900 1792; do
    timeout 100 python bench.py --basis $n --steps 200 --warmup 5 --skip-cpu --skip-65k --skip-e2e --skip-small 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['config']['basis'],'us/term',round(1e3*d['ms_per_step']/24,2))" 2>&1 | tee -a gpurun_out/resident_quick.log
done
