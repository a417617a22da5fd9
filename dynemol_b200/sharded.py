"""Row-sharded operation over the GPUs of one box: one process per GPU (torchrun), H' split by rows,
per-term NCCL reduce-scatter (bra partials) + all-gather (new ket slices) inside the C++ driver.

torch.distributed is plumbing here: it ships the NCCL unique id to the ranks, provides the barrier and the
max-over-ranks of the device-timed region.  Import torch BEFORE dynemol_b200.api in a multi-rank process so
that the library binds the NCCL copy torch already loaded.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def shard_rows(N: int, world: int, rank: int):
    """Uniform contiguous row blocks (the reduce-scatter / all-gather need equal counts)."""
    if N % world != 0:
        raise ValueError(f"N={N} must be divisible by the number of GPUs ({world})")
    m = N // world
    if m % 4 != 0:
        raise ValueError("rows per GPU must be a multiple of 4 (four orbitals per atom)")
    return rank * m, m


def broadcast_unique_id(dist, make_id, device=None):
    """Rank 0 creates the 128-byte NCCL id, every rank receives it (works with the gloo and nccl backends)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if dist.get_rank() == 0:
        raw = make_id()
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().tolist())


def init_sharded(N: int, dist, local_rank: int):
    """Create this rank's context (rows rank*N/P ..) and join the NCCL communicator."""
    import torch
    from dynemol_b200 import api
    world, rank = dist.get_world_size(), dist.get_rank()
    row0, m = shard_rows(N, world, rank)
    P = api.Propagator(N, device=local_rank, row0=row0, n_rows=m)
    dev = torch.device("cuda", local_rank)
    uid = broadcast_unique_id(dist, api.comm_unique_id, device=dev)
    P.comm_init(rank, world, uid)
    if world <= 8 and os.environ.get("DYNEMOL_B200_P2P", "1") != "0":
        # fused exchange over NVLink peer memory: all-gather the 64-byte CUDA-IPC handles, map the peers.  If any rank
        # cannot export or map (no peer access, IPC disabled in the container), EVERY rank falls back to NCCL.
        ok = 1
        try:
            mine = torch.tensor(list(P.comm_p2p_handle()), dtype=torch.uint8, device=dev)
        except api.DynemolB200Error as e:
            ok = 0; mine = torch.zeros(64, dtype=torch.uint8, device=dev); err = e
        allh = torch.empty(64 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            try:
                P.comm_p2p_open(bytes(allh.cpu().tolist()))
            except api.DynemolB200Error as e:
                ok = 0; err = e
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use_p2p = int(flag.item()) == 1
        if not use_p2p:
            P.comm_p2p_enable(False)
            if rank == 0:
                print("dynemol_b200: peer-memory exchange unavailable, using NCCL collectives", flush=True)
        dist.barrier()
    return P, row0, m


def nvlink_counters_kib(index: int):
    """(tx_KiB, rx_KiB) summed over the NVLink links of GPU `index` (nvidia-smi nvlink -gt d), or None."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        tx = sum(int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
        rx = sum(int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
        return (tx, rx) if (tx or rx) else None
    except Exception:
        return None


def synthetic_packets(N: int):
    """el on orbitals 0..63, hole on 64..127; unnormalised S-free variant for the throughput configs."""
    w = 64
    C = np.zeros((N, 2))
    C[0:w, 0] = np.random.default_rng(42).normal(size=w)
    C[w:2 * w, 1] = np.random.default_rng(43).normal(size=w)
    C /= np.linalg.norm(C, axis=0)
    Psi = np.asfortranarray(C.astype(np.complex128))
    return Psi, Psi.copy(order="F")


def bench_main(args):
    import faulthandler
    faulthandler.dump_traceback_later(1500, exit=True)              # never hang a GPU box on a lost collective
    import torch
    import torch.distributed as dist
    from dynemol_b200 import synthetic as syn
    import bench as B

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    N = args.basis or 65536
    P, row0, m = init_sharded(N, dist, local_rank)
    t0 = time.time()
    fill_rows(P, N, row0, m, dev)
    gen_s = time.time() - t0
    bra, ket = synthetic_packets(N)
    P.set_packets(bra, ket)
    tau = B.pick_tau(N)
    info = P.info()

    for _ in range(max(args.warmup, 3)):
        P.run_terms(tau, B.TERMS_PER_STEP)
    nv0 = nvlink_counters_kib(local_rank) if rank == 0 else None      # BEFORE the barrier: the subprocess must not delay rank 0's launches
    dist.barrier(); torch.cuda.synchronize(dev)
    l0 = P.launch_count()
    with B.ClockSampler(local_rank) as cs:
        ms, _ = P.run_terms(tau, B.TERMS_PER_STEP * args.steps)       # CUDA events on the launching stream
        torch.cuda.synchronize(dev)
    dist.barrier()
    nv1 = nvlink_counters_kib(local_rank) if rank == 0 else None
    launches = P.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                          # max over ranks of the device-timed region
    ms_max = float(t.item())
    n_terms = B.TERMS_PER_STEP * args.steps
    value = n_terms / (ms_max * 1e-3)
    clocks = cs.result()

    # end-to-end through the host-buffer API: packets from host memory each call, result read back
    e2e_terms = B.TERMS_PER_STEP * 32          # ~ one nuclear step of dt = 0.5 fs with the single-expansion Chebyshev propagator
    dist.barrier(); torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    P.set_packets(bra, ket)
    P.run_terms(tau, e2e_terms)
    ob, ok = P.get_packets()
    torch.cuda.synchronize(dev); dist.barrier()
    t_e2e = time.perf_counter() - t1
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)

    ref1 = None
    if rank == 0 and not args.no_ref1:
        ref1 = single_gpu_same_workload(args, N, dev, tau, bra, ket)
    dist.barrier()
    P.close()
    # driver-visible correctness of the sharded path: a short checked nuclear step on the same ranks (oracle on rank 0)
    parity = None if args.no_parity else B.sharded_parity_check(dist, local_rank, dev)       # bench.py: the oracle is its checker

    if rank == 0:
        peak, peak_src = B.measured_peak_gbs()
        per_gpu_bytes = 8.0 * m * N
        achieved = per_gpu_bytes * n_terms / (ms_max * 1e-3) / 1e9
        line = {"metric": B.METRIC, "value": round(value, 2), "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": B.workload_name(N, world), "gpus": "%dxB200" % world,
                           "exchange_per_term": "fused NVLink peer-memory exchange (reduce-scatter by peer loads, all-gather by peer stores) inside the epilogue kernel"
                               if info.get("p2p") else "NCCL reduce-scatter(bra)+all-gather(ket)",
                           "exchange": "p2p-fused" if info.get("p2p") else "nccl",
                           "basis": N, "rows_per_gpu": m, "terms_per_step": B.TERMS_PER_STEP,
                           "l2": "inputs larger than L2 (%.2f GB of H' per GPU per pass)" % (per_gpu_bytes / 1e9),
                           "operator": "Hueckel h + dense decaying tail, surrogate for S^-1 h (SURVEY.md 8d)", "grid": info["grid"], "gen_s": round(gen_s, 1)},
                "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                             "traffic": None, "kernel": "whole term incl. collectives, per GPU", "peak_source": peak_src,
                             "alg_bytes_per_launch": per_gpu_bytes},
                "e2e": {"value": round(e2e_terms / float(te.item()), 2), "unit": B.UNIT, "h2d_bytes_per_step": int(2 * 2 * 16 * N),
                        "d2h_bytes_per_step": int(2 * 2 * 16 * N), "call": "set_packets(host)+%d terms+get_packets(host); H' shards resident" % e2e_terms},
                "gpu_launches": int(launches), "clocks": clocks, "single_gpu_same_workload": ref1, "parity_check": parity}
        if nv0 and nv1:
            # NVLink traffic of rank 0's GPU over the timed region (driver counters, all links): the algorithm moves, per
            # term and GPU, (P-1)/P * N * 32 B of bra partials in (peer loads) and as many ket bytes out (peer stores)
            line["nvlink"] = {"tx_kib_per_term": round((nv1[0] - nv0[0]) / n_terms, 1), "rx_kib_per_term": round((nv1[1] - nv0[1]) / n_terms, 1),
                              "algorithmic_kib_per_term_each_way": round((world - 1) / world * N * 32 / 1024.0, 1), "source": "nvidia-smi nvlink -gt d, GPU of rank 0"}
        if ref1 and ref1.get("value"):
            line["speedup_same_workload"] = round(value / ref1["value"], 3)
            line["efficiency_same_workload"] = round(value / ref1["value"] / world, 4)
        print(json.dumps(line))
    dist.destroy_process_group()


def fill_rows(P, N, row0, m, dev, blk=4096):
    """Generate global rows row0..row0+m of the surrogate operator block by block and copy them into P's shard."""
    import torch
    from dynemol_b200 import synthetic as syn
    for r in range(0, m, blk):
        n = min(blk, m - r)
        rows = syn.make_h_shard_colmajor_torch(N, row0 + r, n, dev)     # (N, n) = column-major n x N block
        torch.cuda.synchronize(dev)
        P.upload_hprime_rows_device(rows.data_ptr(), n, r, n)
        del rows
    torch.cuda.empty_cache()


def single_gpu_same_workload(args, N, dev, tau, bra, ket):
    """Strong-scaling denominator measured in the same run: the full N x N operator on rank 0's GPU alone."""
    import torch
    from dynemol_b200 import api
    import bench as B
    free, _ = torch.cuda.mem_get_info(dev)
    need = 8.0 * N * N * 1.05 + 6e9
    if free < need:
        return {"skipped": "not enough free HBM for the full operator (%.0f GB needed, %.0f free)" % (need / 1e9, free / 1e9)}
    P1 = api.Propagator(N, device=dev.index)
    fill_rows(P1, N, 0, N, dev)
    P1.set_packets(bra, ket)
    for _ in range(2):
        P1.run_terms(tau, B.TERMS_PER_STEP)
    steps = max(2, min(args.steps, 10))
    ms, _ = P1.run_terms(tau, B.TERMS_PER_STEP * steps)
    P1.close()
    return {"n_gpus": 1, "value": round(B.TERMS_PER_STEP * steps / (ms * 1e-3), 2), "unit": B.UNIT, "steps": steps}


def config5_main(args):
    """BASELINE config 5: N ~ 30k basis (dye-TiO2-interface size), Taylor vs Chebyshev propagator, 8 x B200.
    H' = S^-1 h is formed on rank 0 from the synthetic EHT S, h at that size (cuSOLVER), its row blocks are scattered to the
    ranks, and ONE nuclear step is propagated with each propagator of the library through the row-sharded path (sharded
    Lanczos bounds included): passes over H' and seconds per nuclear step.  The Taylor series (the reference's shipped
    propagator, Taylor.f) needs tau * rho(H') ~ 1 per series, so it is timed on a slice of the step (--config5-taylor-frac)
    and scaled linearly in dt (every sub-step costs the same); both Chebyshev modes run the full 0.5 fs."""
    import faulthandler
    faulthandler.dump_traceback_later(1500, exit=True)
    import torch
    import torch.distributed as dist
    from dynemol_b200 import api, synthetic as syn
    import bench as B
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    N = args.basis or 30720
    row0, m = shard_rows(N, world, rank)
    t0 = time.time()
    packets = torch.empty((2, N, 2), dtype=torch.float64, device=dev)       # [bra|ket][N][el,hl] (real packets)
    blocks = None
    form_s = None
    if rank == 0:
        S, h, _ = syn.make_S_h_torch(N, dev)
        w = 64
        C = torch.zeros((N, 2), dtype=torch.float64, device=dev)
        C[0:w, 0] = torch.tensor(np.random.default_rng(42).normal(size=w), device=dev)
        C[w:2 * w, 1] = torch.tensor(np.random.default_rng(43).normal(size=w), device=dev)
        SC = S @ C
        nrm = torch.sqrt((C * SC).sum(0))
        packets[0] = SC / nrm; packets[1] = C / nrm
        Pf = api.Propagator(N, device=local_rank)
        torch.cuda.synchronize(dev); t1 = time.time()
        Pf.form_hprime_device(S.data_ptr(), N, h.data_ptr(), N)
        form_s = time.time() - t1
        del S, h, SC, C
        torch.cuda.empty_cache()
        blocks = []
        for r in range(world):
            blk = torch.empty((N, m), dtype=torch.float64, device=dev)      # column-major m x N block of rows r*m ..
            Pf.download_hprime_rows_device(blk.data_ptr(), m, r * m, m)
            blocks.append(blk)
        Pf.close()
        torch.cuda.empty_cache()
    mine = torch.empty((N, m), dtype=torch.float64, device=dev)
    dist.scatter(mine, blocks, src=0)
    dist.broadcast(packets, src=0)
    del blocks
    P, row0, m = init_sharded(N, dist, local_rank)
    P.upload_hprime_rows_device(mine.data_ptr(), m, 0, m)
    del mine
    torch.cuda.empty_cache()
    pk = packets.cpu().numpy()
    Psi_bra = np.asfortranarray(pk[0].astype(np.complex128)); Psi_ket = np.asfortranarray(pk[1].astype(np.complex128))
    build_s = time.time() - t0
    dt = 5e-4
    frac = args.config5_taylor_frac
    res = {}
    P.set_packets(Psi_bra, Psi_ket)
    torch.cuda.synchronize(dev); dist.barrier(); t1 = time.perf_counter()
    lo, hi = P.estimate_spectral_bounds(24, 0.05)
    torch.cuda.synchronize(dev); dist.barrier(); lanczos_s = time.perf_counter() - t1
    for name, mode, dtm in (("chebyshev_single_expansion", api.MODE_CHEBYSHEV_FULL, dt), ("chebyshev_order25_chain", api.MODE_CHEBYSHEV, dt),
                            ("taylor", api.MODE_TAYLOR, dt * frac)):
        tau_max = dtm / api.H_BAR
        P.set_packets(Psi_bra, Psi_ket)
        save, _ = P.propagate(0.0, dtm, tau_max, mode=mode)                 # first step finds tau (untimed)
        P.set_packets(Psi_bra, Psi_ket)
        torch.cuda.synchronize(dev); dist.barrier(); t1 = time.perf_counter()
        save, traces = P.propagate(0.0, dtm, np.minimum(tau_max, 1.15 * save), mode=mode)
        torch.cuda.synchronize(dev); dist.barrier(); el = time.perf_counter() - t1
        te = torch.tensor([el], dtype=torch.float64, device=dev); dist.all_reduce(te, op=dist.ReduceOp.MAX)
        bra, ket = P.get_packets()
        terms = int(P.info()["passes_last"])
        scale = dt / dtm
        res[name] = {"dt_ps_timed": dtm, "terms_timed": terms, "s_timed": round(float(te.item()), 4),
                     "terms_per_0.5fs_step": int(round(terms * scale)), "s_per_0.5fs_step": round(float(te.item()) * scale, 4),
                     "extrapolated": bool(scale != 1.0), "terms_per_s": round(terms / float(te.item()), 1),
                     "norm_el": float(abs(np.vdot(bra[:, 0], ket[:, 0]))), "norm_hl": float(abs(np.vdot(bra[:, 1], ket[:, 1])))}
    if rank == 0:
        line = {"metric": "nuclear step (dt = 0.5 fs) of the el+hole propagator, Taylor vs Chebyshev", "unit": "s per nuclear step", "n_gpus": world,
                "config": {"workload": "BASELINE config 5: synthetic EHT Hamiltonian N=%d basis (dye-TiO2-interface size), H' = S^-1 h formed on one GPU and row-sharded over %d GPUs" % (N, world),
                           "basis": N, "rows_per_gpu": m, "exchange": "p2p-fused" if P.info()["p2p"] else "nccl",
                           "spectral_interval_eV": [lo, hi], "R_dE_tau": round(0.5 * (hi - lo) * dt / api.H_BAR, 1)},
                "propagators": res, "lanczos_24_steps_s": round(lanczos_s, 4), "form_hprime_s": None if form_s is None else round(form_s, 3),
                "build_s": round(build_s, 1), "dtype": "f64", "data": "synthetic"}
        print(json.dumps(line))
    P.close()
    dist.destroy_process_group()
