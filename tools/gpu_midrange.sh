#!/bin/bash
# Mid-size operators through the launch-per-term path with shorter panels (build variants of the same kernels).
mkdir -p gpurun_out
export DYNEMOL_B200_SERIES=term
for lib in libdynemol_b200.so libdyb_sr128_tc4.so libdyb_sr64_tc4.so; do
  export DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/$lib
  for n in ${SIZES:-2048 3000 4096 6144 8192}; do
    timeout 200 python bench.py --basis $n --steps 60 --warmup 5 --skip-cpu --skip-65k --skip-e2e 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print(os.path.basename(os.environ['DYNEMOL_B200_LIB']),'N',d['config']['basis'],'grid',d['config']['grid'],'tiles',d['config']['tiles'],'us/term',round(1e3*d['ms_per_step']/24,2),'kernel_us',r['kernel_avg_us'])" 2>&1 | tee -a gpurun_out/midrange_ab.log
  done
done
