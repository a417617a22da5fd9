#!/bin/bash
# memory-system ceiling of the TMA streaming pattern: the same kernel with the arithmetic removed
for v in "" variant_nomath.so; do
  if [ -n "$v" ]; then export DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/$v; else unset DYNEMOL_B200_LIB; fi
  for i in 1 2; do
  timeout 300 python bench.py --steps 60 --warmup 3 --skip-cpu --skip-e2e --skip-65k 2>&1 | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print('${v:-default}','us/term',round(1e3*d['ms_per_step']/24,2),'kernel_us',r['kernel_avg_us'],'GB/s',r['achieved'],'clk',d['clocks']['sm_mhz'],d['clocks']['reasons'])"
  done
done
NCU=/usr/local/cuda/bin/ncu
export DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/variant_nomath.so
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:dual_matvec_tma -s 30 -c 4 --csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-65k 2>/dev/null | grep -E "dual_matvec" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | head -8
