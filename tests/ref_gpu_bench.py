#!/usr/bin/env python
"""tests/ref_gpu_bench.py -- side measurement (lives under tests/ because it executes oracle/_ref: checker code, never the product), NOT part of bench.py's contract: the reference's OWN GPU kernels
(oracle/_ref/libref_taylor_gpu.so = Taylor_gpu.cpp + dzgemv_kernels.cu compiled in place for sm_100a) timed on the
same B200 next to the product.

  1. kblas_dzgemv2_async (dzgemv_kernels.cu, the H'·psi kernel of the reference: real matrix x one complex vector),
     'N' and 'T', N = 16384 and 4096: achieved GB/s.  One el+hole term of the reference = 2 'N' + 2 'T' such passes
     (bra and ket of each particle, on two MPI ranks in the reference); the product needs ONE pass for all four.
  2. One nuclear step through the Fortran symbol both libraries export (propagationelhl_gpucaller_), same host inputs,
     electron then hole (the reference runs them on two ranks / two GPUs; here back to back on one): wall seconds.
Writes one JSON object to stdout and to gpurun_out/ref_gpu_bench.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
H_BAR = 6.58264e-4


def main():
    import torch
    import oracle
    from dynemol_b200 import api, synthetic as syn
    assert oracle.ref_gpu_available(), "oracle/_ref/libref_taylor_gpu.so missing or no GPU"
    ref = oracle._tgpu()
    dev = torch.device("cuda", 0)
    out = {"device": torch.cuda.get_device_name(0)}

    # ---- 1. the reference's dzgemv kernel, in place on a resident matrix
    ref.kblas_dzgemv2_async.argtypes = [C.c_char, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    ref.kblas_dzgemv2_async.restype = C.c_int
    kern = {}
    for N in (4096, 16384):
        ld = ((N + 31) // 32) * 32
        A = torch.randn((N, ld), dtype=torch.float64, device=dev)           # column-major ld x N seen from the kernel
        x = torch.randn((ld, 2), dtype=torch.float64, device=dev)           # interleaved complex
        y = torch.zeros((ld, 2), dtype=torch.float64, device=dev)
        s = torch.cuda.Stream()
        for trans in (b"n", b"t"):
            with torch.cuda.stream(s):
                for _ in range(3):
                    ref.kblas_dzgemv2_async(trans, N, 1.0, A.data_ptr(), ld, x.data_ptr(), 0.0, y.data_ptr(), s.cuda_stream)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                reps = 20
                e0.record(s)
                for _ in range(reps):
                    ref.kblas_dzgemv2_async(trans, N, 1.0, A.data_ptr(), ld, x.data_ptr(), 0.0, y.data_ptr(), s.cuda_stream)
                e1.record(s)
            s.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            kern[f"N{N}_{trans.decode()}"] = {"us": round(us, 1), "GBs": round(8.0 * N * N / (us * 1e-6) / 1e9, 1)}
        # the product on the same size: one dual product serving el+hole, bra+ket
        P = api.Propagator(N)
        P.upload_hprime_device(A.data_ptr(), ld)
        rng = np.random.default_rng(1)
        v = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
        P.set_packets(v, v)
        P.set_series_kernel("term")
        P.run_terms(1e-4, 24)
        ms, kms = P.run_terms(1e-4, 48, per_kernel=True)
        kern[f"N{N}_product_dual_pass"] = {"us": round(kms * 1e3 / 48, 1), "GBs": round(8.0 * N * N / (kms * 1e-3 / 48) / 1e9, 1),
                                           "us_per_term_with_epilogue": round(ms * 1e3 / 48, 1)}
        tn, tt = kern[f"N{N}_n"]["us"], kern[f"N{N}_t"]["us"]
        kern[f"N{N}_reference_us_per_elhl_term_lower_bound"] = round(2 * tn + 2 * tt, 1)
        P.close()
        del A, x, y
        torch.cuda.empty_cache()
    out["kernels"] = kern

    # ---- 2. one nuclear step through the shared Fortran symbol
    N, dt = 4096, 1e-6
    w = syn.make_workload(N)
    tau0 = dt / H_BAR
    step = {"N": N, "dt_ps": dt}
    for name in ("warmup", "timed"):
        t0 = time.time()
        r = [oracle.ref_gpu_propagationelhl(w.S, w.h, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0) for p in range(2)]
        t1 = time.time()
        o = [api.legacy_propagationelhl(w.S, w.h, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0) for p in range(2)]
        t2 = time.time()
        ob = api.legacy_propagationelhl(w.S, w.h, w.Psi_bra.copy(), w.Psi_ket.copy(), 0.0, dt, tau0)
        t3 = time.time()
        step[name] = {"reference_gpu_s": round(t1 - t0, 4), "product_two_calls_s": round(t2 - t1, 4), "product_batched_call_s": round(t3 - t2, 4)}
    err = max(np.abs(o[p]["PSI_ket"] - r[p][3]).max() / np.abs(r[p][3]).max() for p in range(2))
    step["max_rel_diff_ket"] = float(err)
    step["save_tau"] = {"reference": [r[0][4], r[1][4]], "product": [float(o[0]["save_tau"][0]), float(o[1]["save_tau"][0])]}
    out["nuclear_step"] = step
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_gpu_bench.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
