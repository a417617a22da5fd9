#!/bin/bash
# Streamed 2-D block series kernel (mid-size operators): parity, then A/B against launch-per-term.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_series_kernels.py -x -q -m gpu -k "blocked" --durations=6 2>&1 | tail -25 | tee gpurun_out/blocked_tests.log
for kind in ${KINDS:-term blocked}; do
  export DYNEMOL_B200_SERIES=$kind
  for n in ${SIZES:-2048 3000 4096 6144}; do
    timeout 200 python bench.py --basis $n --steps 60 --warmup 5 --skip-cpu --skip-65k --skip-e2e 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print('series',os.environ['DYNEMOL_B200_SERIES'],'N',d['config']['basis'],'us/term',round(1e3*d['ms_per_step']/24,2),'value',round(d['value'],1),'launches',d['gpu_launches'])" 2>&1 | tee -a gpurun_out/blocked_ab.log
  done
done
