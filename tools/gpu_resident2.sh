#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chebyshev.py tests/test_gpu_poststep.py tests/test_gpu_driver.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/resident_tests.log
SIZES="128 512 900 1792" bash tools/gpu_resident_prof.sh
export DYNEMOL_B200_SERIES=auto
for n in 128 512 900 1792; do
    timeout 200 python bench.py --basis $n --steps 100 --warmup 5 --skip-cpu --skip-65k --skip-e2e 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print('series',os.environ['DYNEMOL_B200_SERIES'],'N',d['config']['basis'],'us/term',round(1e3*d['ms_per_step']/24,2),'value',round(d['value'],1),'launches',d['gpu_launches'])" 2>&1 | tee -a gpurun_out/resident_ab.log
done
