#!/bin/bash
# multi-GPU check: sharded parity test + sharded bench.  usage: gpu_multi.sh NGPU BASIS STEPS
NG=${1:-2}; BASIS=${2:-32768}; STEPS=${3:-10}
mkdir -p gpurun_out
nvidia-smi -L
echo "=== sharded parity"; timeout 600 python -m pytest tests/test_gpu_sharded.py -q -x 2>&1 | tail -15
echo "=== sharded bench N=$BASIS on $NG GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577 \
   bench.py --gpus $NG --steps $STEPS --warmup 3 --basis $BASIS 2>&1 | grep -v "^W\|^\[W\|^$" | tail -8 | tee gpurun_out/bench_multi_${NG}_${BASIS}.json
