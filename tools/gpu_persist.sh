#!/bin/bash
# A/B of the one-launch-per-series kernel against the launch-per-term path: parity first, then timing.
mkdir -p gpurun_out
export DYNEMOL_B200_SERIES=stream
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --durations=5 2>&1 | tail -15 | tee gpurun_out/persist_tests.log
for pers in ${PERS:-stream}; do
  export DYNEMOL_B200_SERIES=$pers
  for n in ${SIZES:-512 2048 4096 16384}; do
    timeout 200 python bench.py --basis $n --steps 40 --warmup 4 --skip-cpu --skip-65k --skip-e2e 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print('persistent',os.environ['DYNEMOL_B200_SERIES'],'N',d['config']['basis'],'grid',d['config']['grid'],'us/term',round(1e3*d['ms_per_step']/24,2),'value',round(d['value'],1),'launches',d['gpu_launches'])" 2>&1 | tee -a gpurun_out/persist_ab.log
  done
done
