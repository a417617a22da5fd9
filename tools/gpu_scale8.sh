#!/bin/bash
# 8-GPU box: sharded parity at world 2 and 8, then strong scaling of N=65536 at 8, 4, 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "=== sharded parity"; timeout 400 python -m pytest tests/test_gpu_sharded.py -q -x 2>&1 | tail -5
for NG in 8 4 2; do
  echo "=== N=65536 on $NG GPUs"
  EXTRA=""; if [ $NG != 8 ]; then EXTRA="--no-ref1"; fi
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2957$NG \
     bench.py --gpus $NG --steps 20 --warmup 3 $EXTRA 2>&1 | grep "^{" | tail -1 | tee gpurun_out/bench_scale_${NG}.json
done
