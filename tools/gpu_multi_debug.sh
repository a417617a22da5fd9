#!/bin/bash
export DYB_TEST_VERBOSE=1 DYB_TEST_WATCHDOG=60 DYB_TEST_N=${1:-1024}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/sharded_worker.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -60
