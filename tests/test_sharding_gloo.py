"""CPU suite, world_size 2 over gloo: the host-side logic of the row-sharded (N>1 GPU) path --
row partition, shipping the NCCL unique id, and the exchange protocol itself (bra partials are
reduce-scattered, new ket slices all-gathered, per-rank scalars combined in rank order), emulated
with numpy on the same synthetic operator.  The CUDA kernels are not involved here."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from dynemol_b200 import sharded, synthetic as syn
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the unique id reaches every rank unchanged
        secret = bytes((7 * i + 3) % 256 for i in range(128))
        got = sharded.broadcast_unique_id(dist, lambda: secret)
        assert got == secret

        # 2. uniform row partition
        N = 64
        row0, m = sharded.shard_rows(N, world, rank)
        assert m == N // world and row0 == rank * m

        # 3. one series term under the sharded protocol == the unsharded term
        w = syn.make_workload(N)
        H = np.linalg.solve(w.S, w.h)                       # H' = S^-1 h
        xk = w.Psi_ket.copy(); xb = w.Psi_bra.copy()
        alpha = -0.25j
        rows = slice(row0, row0 + m)
        ket_local = alpha * (H[rows, :] @ xk)               # complete for the owned rows
        bra_partial = H[rows, :].T @ xb[rows, :]            # full length, partial over the owned rows
        t = torch.tensor(np.stack([bra_partial.real, bra_partial.imag]))
        dist.all_reduce(t)                                  # gloo has no reduce_scatter: all-reduce + slice
        bra_local = alpha * (t[0].numpy() + 1j * t[1].numpy())[rows, :]
        gathered = [torch.zeros(2, m, 2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor(np.stack([ket_local.real, ket_local.imag])))
        ket_full = np.concatenate([g[0].numpy() + 1j * g[1].numpy() for g in gathered], axis=0)
        assert np.allclose(ket_full, alpha * (H @ xk), rtol=1e-12, atol=1e-14)
        assert np.allclose(bra_local, (alpha * (H.T @ xb))[rows, :], rtol=1e-12, atol=1e-14)

        # 4. scalars: per-rank partial dot products / maxima combined in rank order give the global ones
        part = np.array([np.abs(ket_local).max(), np.vdot(bra_local[:, 0], ket_local[:, 0]).real])
        allp = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allp, torch.tensor(part))
        mx = max(float(p[0]) for p in allp); dot = sum(float(p[1]) for p in allp)
        full_k = alpha * (H @ xk); full_b = alpha * (H.T @ xb)
        assert abs(mx - np.abs(full_k).max()) < 1e-14
        assert abs(dot - np.vdot(full_b[:, 0], full_k[:, 0]).real) < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_rows_rejects_ragged():
    sys.path.insert(0, ROOT)
    from dynemol_b200 import sharded
    with pytest.raises(ValueError):
        sharded.shard_rows(100, 8, 0)
    assert sharded.shard_rows(65536, 8, 3) == (3 * 8192, 8192)
