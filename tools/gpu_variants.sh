#!/bin/bash
# A/B of tuning builds of the dual-product kernel (same box, back to back)
mkdir -p gpurun_out
for v in "" $@; do
  if [ -n "$v" ]; then export DYNEMOL_B200_LIB=$PWD/dynemol_b200/lib/$v; else unset DYNEMOL_B200_LIB; fi
  echo "=== variant: ${v:-default}"
  timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "dual_matvec or deterministic or golden" 2>&1 | tail -2
  timeout 300 python bench.py --steps 40 --warmup 3 --skip-cpu --skip-e2e --skip-65k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value',d['value'],'kernel_us',r['kernel_avg_us'],'GB/s',r['achieved'],'frac',r['frac'],'clk',d['clocks']['sm_mhz'],d['clocks']['reasons'])"
done
