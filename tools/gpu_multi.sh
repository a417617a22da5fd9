#!/bin/bash
# multi-GPU check: sharded parity (P2P and NCCL exchange) + sharded bench both ways.  usage: gpu_multi.sh NGPU BASIS STEPS
NG=${1:-2}; BASIS=${2:-32768}; STEPS=${3:-10}
mkdir -p gpurun_out
echo "=== sharded parity"; timeout 500 python -m pytest tests/test_gpu_sharded.py -q -x 2>&1 | tail -6
for P2P in 1 0; do
  echo "=== sharded bench N=$BASIS on $NG GPUs, P2P=$P2P"
  DYNEMOL_B200_P2P=$P2P timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2957$P2P \
     bench.py --gpus $NG --steps $STEPS --warmup 3 --basis $BASIS --no-ref1 2>&1 | grep "^{\|rror\|timeout" | tail -3 | tee gpurun_out/bench_multi_${NG}_${BASIS}_p2p$P2P.json | python -c "
import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print('value',d['value'],'ms/step',d['ms_per_step'],'exchange',d['config'].get('exchange'),'frac',d['roofline']['frac'],'clk',d['clocks']['sm_mhz'])
    except Exception as e: print(l[:300])"
done
