"""GPU suite, needs >= 2 GPUs (skipped on a 1-GPU box): the row-sharded path, one process per GPU over NCCL,
against the CPU oracle: identical decision traces, wavepackets within 1e-10."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("p2p", [1, 0])          # 1: fused peer-memory exchange (NVLink P2P), 0: NCCL collectives
@pytest.mark.parametrize("world", [2, 8])
def test_sharded_propagation_matches_oracle(world, p2p):
    from dynemol_b200 import api
    if api.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import oracle
    oracle.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "sharded_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=dict(os.environ, DYB_TEST_WATCHDOG="200", DYNEMOL_B200_P2P=str(p2p)))
    lines = [l for l in res.stdout.splitlines() if l.startswith("SHARDED_RESULT ")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1][len("SHARDED_RESULT "):])
    assert out["ok"], out
    assert out["p2p"] == p2p
