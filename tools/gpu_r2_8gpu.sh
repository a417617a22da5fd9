#!/bin/bash
# Round 2, 8-GPU box: team + sharded parity tests, the SCALE line at 8 GPUs, BASELINE config 5 (one process per GPU and
# single-process team through the primary symbol).  gpurun --gpus 8 --timeout 1200 -- bash tools/gpu_r2_8gpu.sh
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_team.py tests/test_gpu_sharded.py -m gpu -x -q -s --durations=8 > gpurun_out/r2f_tests8.log 2>&1; tail -12 gpurun_out/r2f_tests8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2f_bench8.log 2>&1; tail -1 gpurun_out/r2f_bench8.log | cut -c1-3000
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config5 > gpurun_out/r2f_config5.log 2>&1; tail -1 gpurun_out/r2f_config5.log | cut -c1-3000
timeout 400 python bench.py --team 8 --basis 30720 > gpurun_out/r2f_team8.log 2>&1; tail -1 gpurun_out/r2f_team8.log | cut -c1-2500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --no-parity > gpurun_out/r2f_bench4.log 2>&1; tail -1 gpurun_out/r2f_bench4.log | cut -c1-600
