#!/bin/bash
# Round 2, 8-GPU box, second pass: team tests (distributed formation), SCALE lines at 8 and 4 GPUs (sampling-order fix),
# the primary symbol on a team of 8 at N=30720 (BASELINE config 5 size) and N=16384.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_team.py -m gpu -x -q -s --durations=5 > gpurun_out/r2h_tests8.log 2>&1; tail -8 gpurun_out/r2h_tests8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2h_bench8.log 2>&1; tail -1 gpurun_out/r2h_bench8.log | cut -c1-400; tail -1 gpurun_out/r2h_bench8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','speedup_same_workload','efficiency_same_workload','nvlink')}, d['e2e']['value'], d['parity_check']['ok'])"
timeout 300 python bench.py --team 8 --basis 30720 > gpurun_out/r2h_team8.log 2>&1; tail -1 gpurun_out/r2h_team8.log | cut -c1-1600
timeout 300 python bench.py --team 8 --basis 16384 > gpurun_out/r2h_team8_16k.log 2>&1; tail -1 gpurun_out/r2h_team8_16k.log | cut -c1-1600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --no-parity > gpurun_out/r2h_bench4.log 2>&1; tail -1 gpurun_out/r2h_bench4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: d.get(k) for k in ('value','speedup_same_workload','efficiency_same_workload','nvlink')}, d['e2e']['value'])"
