#!/usr/bin/env python
"""bench.py -- el+hole series terms/s of the B200-native propagator (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--basis NB]

One "step" = one 24-term el+hole series (series_init + 24 x (dual-product pass over H' + fused
epilogue)), i.e. the work of one Convergence() call of the reference (Taylor.f:132-219) for the
electron and the hole together.  Unit of work = one el+hole term = one pass over H' serving the
four complex right-hand sides: 8*N^2 algorithmic bytes.

N = 1: workload "synthetic EHT Hamiltonian N=16384, el+hole packets" (BASELINE config 3), H' = S^-1 h
formed on the device from the synthetic S, h.  N > 1: N = 65536 row-sharded (BASELINE config 4).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TERMS_PER_STEP = 24
H_BAR = 6.58264e-4
METRIC = "el+hole Chebyshev terms/s"
UNIT = "terms/s"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region (NVML; nvidia-smi as a fallback)."""

    def __init__(self, index=0, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        if nv is not None:
            for k in ("SwPowerCap", "HwSlowdown", "SwThermalSlowdown", "HwThermalSlowdown", "HwPowerBrakeSlowdown",
                      "SyncBoost", "ApplicationsClocksSetting", "DisplayClockSetting"):
                v = getattr(nv, "nvmlClocksEventReason" + k, None) or getattr(nv, "nvmlClocksThrottleReason" + k, None)
                if v is not None:
                    names[int(v)] = k
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if r & bit:
                            self.reasons.add(nm)
                else:
                    import subprocess
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip().split(",")
                    self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                    for nm, v in zip(("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap"), out[2:]):
                        if "Active" in v and "Not" not in v:
                            self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2.0)

    def result(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(N):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(str(N))
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------- workloads
def build_single_gpu(N, device=0, keep_host=False):
    """Synthetic EHT S, h on the device (input generation with torch = plumbing), H' = S^-1 h formed by the
    library (cuSOLVER potrf + potrs), packets on the host."""
    import torch
    from dynemol_b200 import api, synthetic as syn
    dev = torch.device("cuda", device)
    t0 = time.time()
    S, h, meta = syn.make_S_h_torch(N, dev)
    P = api.Propagator(N, device=device)
    torch.cuda.synchronize(dev)
    t1 = time.time()
    P.form_hprime_device(S.data_ptr(), N, h.data_ptr(), N)         # S, h symmetric: row-major == column-major
    t2 = time.time()
    # packets: el on the first 64 orbitals, hole on the next 64; Psi_bra = S C, Psi_ket = C (ElHl_Chebyshev.f:126-129)
    w = 64
    C = np.zeros((N, 2))
    C[0:w, 0] = np.random.default_rng(42).normal(size=w)
    C[w:2 * w, 1] = np.random.default_rng(43).normal(size=w)
    Ct = torch.tensor(C, device=dev)
    SC = S @ Ct
    nrm = torch.sqrt((Ct * SC).sum(0))
    Ct = Ct / nrm; SC = SC / nrm
    Psi_ket = np.asfortranarray(Ct.cpu().numpy().astype(np.complex128))
    Psi_bra = np.asfortranarray(SC.cpu().numpy().astype(np.complex128))
    host = None
    if keep_host:                                  # pinned host copies of S and h for the legacy-symbol end-to-end leg
        Sh = torch.empty((N, N), dtype=torch.float64, pin_memory=True); hh = torch.empty((N, N), dtype=torch.float64, pin_memory=True)
        Sh.copy_(S); hh.copy_(h)
        host = (Sh, hh)
    del S, h, SC, Ct
    torch.cuda.empty_cache()
    return P, Psi_bra, Psi_ket, {"gen_s": t1 - t0, "form_hprime_s": t2 - t1}, host


def pick_tau(N):
    # keeps every 24-term series bounded (|r_k| * ||H'|| < 1 for k >= 2): the same regime Convergence() settles in
    return 1.0e-4


# ----------------------------------------------------------------------------------------------- arms
def run_ours_single(args):
    import torch
    from dynemol_b200 import api
    N = args.basis or 16384
    if N > 32768:
        # forming S^-1 h at this size needs ~3 N^2 doubles + a 6.5e14-flop solve: use the Hueckel surrogate
        from dynemol_b200 import sharded
        P = api.Propagator(N)
        t0 = time.time(); sharded.fill_rows(P, N, 0, N, torch.device("cuda", 0)); build_info = {"gen_s": time.time() - t0, "operator": "Hueckel h + dense decaying tail (surrogate)"}
        Psi_bra, Psi_ket = sharded.synthetic_packets(N)
        args.skip_e2e = True; args.skip_cpu = True; host_Sh = None
    else:
        P, Psi_bra, Psi_ket, build_info, host_Sh = build_single_gpu(N, keep_host=not args.skip_e2e)
    P.set_packets(Psi_bra, Psi_ket)
    if args.kernel == "ldg":
        P.set_kernel(api.KERNEL_LDG)
    tau = pick_tau(N)
    info = P.info()

    for _ in range(max(args.warmup, 3)):
        P.run_terms(tau, TERMS_PER_STEP)
    torch.cuda.synchronize()
    l0 = P.launch_count()
    with ClockSampler(0) as cs:
        P.sync()
        ms, _ = P.run_terms(tau, TERMS_PER_STEP * args.steps)      # CUDA events on the launching stream, inside the library
        P.sync()
    launches = P.launch_count() - l0
    clocks = cs.result()
    n_terms = TERMS_PER_STEP * args.steps
    value = n_terms / (ms * 1e-3)

    # dominant kernel: events around every dual-product launch (separate run so the brackets do not perturb `value`)
    ms2, kms = P.run_terms(tau, TERMS_PER_STEP * min(args.steps, 50), per_kernel=True)
    k_launches = TERMS_PER_STEP * min(args.steps, 50)
    alg_bytes = 8.0 * N * N
    k_avg_s = (kms * 1e-3) / k_launches
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / k_avg_s / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": recorded_traffic(N), "kernel": "dual_matvec_%s_kernel" % args.kernel, "kernel_avg_us": round(k_avg_s * 1e6, 2),
                "alg_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": round(kms / ms2, 4)}

    # e2e: the reference-facing call sequence with HOST buffers (propagation_gpucaller_ semantics, Taylor_gpu.cpp:295-330):
    # H' and packets start in pinned host memory, H2D + full propagation of one nuclear step + D2H inside the timed region.
    # e2e = the PRIMARY plugin symbol with host buffers ("make the primary symbol the e2e"); the round-1
    # definition (native API, H' uploaded ready-made: propagation_gpucaller_ semantics) is kept beside it as e2e.native_api
    if args.skip_e2e:
        e2e = None
    else:
        native = run_e2e(args, P, N, Psi_bra, Psi_ket)
        e2e = None
        if host_Sh is not None:
            try:
                leg = run_e2e_legacy(args, N, Psi_bra, Psi_ket, host_Sh)
                m25 = leg["modes"]["chebyshev25"]
                e2e = {"value": m25["terms_per_s"], "unit": UNIT, "h2d_bytes_per_step": leg["h2d_bytes_per_step"],
                       "d2h_bytes_per_step": leg["d2h_bytes_per_step"], "call": leg["call"] + ", DYNEMOL_B200_MODE=chebyshev25",
                       "terms_per_call": m25["terms_per_call"], "s_per_call": m25["s_per_call"], "nuclear_steps_per_s": m25["nuclear_steps_per_s"],
                       "single_expansion": leg["modes"]["chebyshev"], "native_api": native,
                       "note": "series terms only are counted; formation (O(N^3)), 24 Lanczos passes and all copies are timed"}
            except Exception as ex:     # never lose the headline line
                e2e = dict(native, legacy_symbol_error=repr(ex)[:300])
            host_Sh = None
        if e2e is None:
            e2e = native
    cpu = None if args.skip_cpu else cpu_baseline(np.asfortranarray(P.download_hprime()), Psi_bra, Psi_ket, tau, budget_s=args.cpu_budget)

    small = None
    if N == 16384 and not args.skip_small:
        small = small_operator(args)

    mid = None
    if N == 16384 and not args.skip_small:
        mid = mid_operator(args)

    n65536 = None
    if N == 16384 and not args.skip_65k:
        n65536 = single_gpu_65536(P, args)

    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N, 1), "basis": N, "gpus": "1xB200",
                       "terms_per_step": TERMS_PER_STEP, "l2": "inputs larger than L2 (H' = %.2f GB per pass)" % (alg_bytes / 1e9),
                       "kernel_variant": args.kernel, "grid": info["grid"], "tiles": info["tiles"], "build": build_info},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "n65536_single_gpu": n65536, "small_operator": small, "mid_operator": mid}
    print(json.dumps(line))


def small_operator(args, N=900):
    """Side measurement at the QM-region size of BASELINE config 2 (heptazine + water droplet, N ~ 900):
    the same 24-term series through the launch-per-term path and through the shared-memory-resident series kernel
    (csrc/resident.cuh), which is what the library selects by itself at this size."""
    import torch
    try:
        P, bra, ket, _, _ = build_single_gpu(N)
        P.set_packets(bra, ket)
        tau = pick_tau(N)
        out = {"basis": N, "unit": UNIT, "steps": 200}
        for kind in ("term", "auto"):
            P.set_series_kernel(kind)
            for _ in range(5):
                P.run_terms(tau, TERMS_PER_STEP)
            ms, _ = P.run_terms(tau, TERMS_PER_STEP * 200)
            key = "per_term" if kind == "term" else ("resident" if P.info()["series_kernel"] == 3 else "auto")
            out[key] = {"value": round(TERMS_PER_STEP * 200 / (ms * 1e-3), 1), "us_per_term": round(ms * 1e3 / (TERMS_PER_STEP * 200), 2)}
            # whole nuclear steps through dyb_propagate (Taylor mode, dt = 0.02 fs, carried-over tau): host logic included
            dt = 2e-5; tau0 = dt / H_BAR
            P.set_packets(bra, ket)
            save, _ = P.propagate(0.0, dt, tau0)
            P.sync(); l0 = P.launch_count(); t0 = time.perf_counter()
            for _ in range(5):
                save, _ = P.propagate(0.0, dt, np.minimum(tau0, 1.15 * save))
            P.sync()
            out[key]["taylor_step_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
            out[key]["passes_per_step"] = P.info()["passes_last"]; out[key]["launches_per_step"] = (P.launch_count() - l0) // 5
        P.close()
        torch.cuda.empty_cache()
        out["config2_100_steps"] = config2_trajectory(N)
        return out
    except Exception as e:      # the headline line must survive a failure of the side measurement
        return {"error": repr(e)[:200]}


def config2_trajectory(N=900, n_steps=100, dt=2e-5):
    """BASELINE config 2 (QM region of the heptazine-in-water example, N ~ 900), the propagator's share of it: 100
    nuclear steps with MOVING nuclei through the PRIMARY symbol propagationelhl2_gpucaller_, called by reference like
    ElHl_Chebyshev_GPU.f:269-272 -- every step uploads that frame's host-built S and h, forms H' = S^-1 h, propagates electron
    and hole (Taylor.f semantics, carried-over tau of ElHl_Chebyshev.f:182-184) and returns H', the packets and AO_bra.
    The frames (S, h of 10 perturbed geometries, reused cyclically) are built before the clock starts: assembling them is
    the reference host's job (overlap_D.f, hamiltonians.f)."""
    from dynemol_b200 import api, synthetic as syn
    try:
        pos, species = syn.lattice(N // 4, 1234 + N)
        frames = [syn.workload_at(syn.perturb_positions(pos, s), species) for s in range(10)]
        _, bra, ket = syn.packets(frames[0][0], N)
        tau_max = dt / H_BAR
        o = api.legacy_propagationelhl(frames[0][0], frames[0][1], bra, ket, 0.0, dt, tau_max, copy_inputs=False)   # step 1, untimed
        save, bra, ket, t, terms = o["save_tau"], o["PSI_bra"], o["PSI_ket"], dt, 0
        t0 = time.perf_counter()
        for step in range(1, n_steps + 1):
            S, h = frames[step % len(frames)]
            o = api.legacy_propagationelhl(S, h, bra, ket, t, t + dt, np.minimum(tau_max, 1.15 * save), copy_inputs=False)
            save, bra, ket = o["save_tau"], o["PSI_bra"], o["PSI_ket"]
            terms += api.legacy_passes_last(); t += dt
        el = time.perf_counter() - t0
        return {"basis": N, "nuclear_steps": n_steps, "dt_ps": dt, "s_total": round(el, 3), "ms_per_nuclear_step": round(el / n_steps * 1e3, 3),
                "terms": int(terms), "terms_per_s": round(terms / el, 1), "norm_el": float(abs(np.vdot(bra[:, 0], ket[:, 0]))),
                "call": "propagationelhl2_gpucaller_ (host S, h of the frame -> H', packets, AO_bra), Taylor.f semantics"}
    except Exception as e:
        return {"error": repr(e)[:200]}
    finally:
        api.gpu_finalize()


def mid_operator(args, sizes=(2048, 4096)):
    """Side measurement in the size range of real Dynemol QM regions (1824 < N <= 6144): the same 24-term series through
    the two-launch path and through the streamed one-launch series kernel (csrc/mid.cuh: H' streamed per term by a TMA ring
    that runs across the terms, barrier-free exchange through epoch-tagged words), which is what the library selects by
    itself there.  H' is a dense surrogate operator generated on the device (only the time per term is measured)."""
    import torch
    from dynemol_b200 import api
    out = {"unit": "us per el+hole term", "series_terms": TERMS_PER_STEP, "steps": 100}
    try:
        for N in sizes:
            g = torch.Generator(device="cuda").manual_seed(N)
            H = torch.randn((N, N), device="cuda", dtype=torch.float64, generator=g) / np.sqrt(N)
            rng = np.random.default_rng(1)
            x = (rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))) / np.sqrt(N)
            P = api.Propagator(N)
            P.upload_hprime_device(H.data_ptr(), N)
            P.set_packets(x, x.conj())
            row = {"hbm_time_of_one_pass_us": round(8.0 * N * N / (measured_peak_gbs()[0] * 1e9) * 1e6, 2)}
            for kind in ("term", "auto"):
                P.set_series_kernel(kind)
                for _ in range(3):
                    P.run_terms(1e-4, TERMS_PER_STEP)
                ms, _ = P.run_terms(1e-4, TERMS_PER_STEP * 100)
                key = "two_launch" if kind == "term" else ("one_launch" if P.info()["series_kernel"] == 5 else "auto")
                row[key] = round(ms * 1e3 / (TERMS_PER_STEP * 100), 2)
            out["N%d" % N] = row
            P.close()
            del H
        torch.cuda.empty_cache()
        return out
    except Exception as e:      # the headline line must survive a failure of the side measurement
        return {"error": repr(e)[:200]}


def single_gpu_65536(P_small, args):
    """The strong-scaling denominator of BASELINE config 4 (N=65536 on ONE GPU, 34 GB of H'), measured in the same
    run so that the --gpus 2/4/8 lines (which use N=65536) can be compared with it."""
    import torch
    from dynemol_b200 import api, sharded
    try:
        P_small.close()
        torch.cuda.empty_cache()
        N = 65536
        free, _ = torch.cuda.mem_get_info()
        if free < 8.0 * N * N * 1.05 + 8e9:
            return {"skipped": "not enough free HBM"}
        P = api.Propagator(N)
        t0 = time.time(); sharded.fill_rows(P, N, 0, N, torch.device("cuda", 0)); gen = time.time() - t0
        bra, ket = sharded.synthetic_packets(N)
        P.set_packets(bra, ket)
        tau = pick_tau(N)
        for _ in range(2):
            P.run_terms(tau, TERMS_PER_STEP)
        steps = 5
        ms, _ = P.run_terms(tau, TERMS_PER_STEP * steps)
        ms2, kms = P.run_terms(tau, TERMS_PER_STEP * 2, per_kernel=True)
        peak, _ = measured_peak_gbs()
        ach = 8.0 * N * N / ((kms * 1e-3) / (TERMS_PER_STEP * 2)) / 1e9
        P.close()
        return {"value": round(TERMS_PER_STEP * steps / (ms * 1e-3), 2), "unit": UNIT, "steps": steps, "basis": N,
                "operator": "Hueckel h + dense decaying tail, surrogate for S^-1 h (SURVEY.md 8d)", "kernel_GBs": round(ach, 1), "frac": round(ach / peak, 4), "gen_s": round(gen, 1)}
    except Exception as e:  # the headline line must survive a failure of the side measurement
        return {"error": repr(e)[:200]}


def run_e2e(args, P, N, Psi_bra, Psi_ket):
    """The reference-facing call sequence with HOST buffers (propagation_gpucaller_ semantics, Taylor_gpu.cpp:295-330):
    H' and the packets start in pinned host memory; H2D, a full propagation of one nuclear step and D2H are inside the
    timed region.  Default: Chebyshev mode, dt = 0.5 fs (BASELINE config 3), tau carried over from the previous step
    like ElHl_Chebyshev.f:182-184; the spectral interval is re-estimated (24 Lanczos passes) inside every call."""
    import torch
    from dynemol_b200 import api
    Hp_host = torch.empty((N, N), dtype=torch.float64, pin_memory=True)
    Hp_np = Hp_host.numpy().T                                       # Fortran-ordered view of the pinned buffer
    Hp_np[...] = P.download_hprime()
    cheb = args.e2e_mode == "cheb"
    dt_e2e = args.e2e_dt if args.e2e_dt > 0 else (5e-4 if cheb else 5e-6)
    tau_max = dt_e2e / H_BAR
    mode = api.MODE_CHEBYSHEV if cheb else api.MODE_TAYLOR

    def one_call(tau):
        P.upload_hprime(Hp_np)                                      # 8 N^2 bytes H2D
        P.set_packets(Psi_bra, Psi_ket)
        if cheb:
            P.estimate_spectral_bounds(24, 0.05)
        save_tau, _ = P.propagate(0.0, dt_e2e, tau, mode=mode)
        P.get_packets()
        return save_tau, P.info()["passes_last"]

    save_tau, _ = one_call(tau_max)                                 # untimed: the first nuclear step finds tau
    e2e_steps = max(1, args.e2e_steps)
    passes = 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, n = one_call(np.minimum(tau_max, 1.15 * save_tau))
        passes += n
    t_e2e = time.perf_counter() - t0
    return {"value": round(passes / t_e2e, 2), "unit": UNIT,
            "h2d_bytes_per_step": int(8 * N * N + 2 * 2 * 16 * N), "d2h_bytes_per_step": int(2 * 2 * 16 * N),
            "call": "upload_hprime(host)+set_packets(host)+%spropagate(%s, dt=%g ps)+get_packets(host)"
                    % ("estimate_spectral_bounds(24)+" if cheb else "", "Chebyshev" if cheb else "Taylor", dt_e2e),
            "terms_per_call": passes // e2e_steps, "s_per_call": round(t_e2e / e2e_steps, 4),
            "note": "series terms only; the 24 Lanczos passes per call are timed but not counted" if cheb else ""}


def run_e2e_legacy(args, N, Psi_bra, Psi_ket, host_Sh):
    """The PRIMARY boundary symbol (batched form): propagationelhl2_gpucaller_(N, S, h, H', AO_bra, AO_ket, PSI_bra,
    PSI_ket, t_init, t_max, tau, save_tau) called by reference exactly like ElHl_Chebyshev_GPU.f:269-272, with the
    caller-side buffers of that routine: S_matrix, h0 and H_prime allocated ONCE and page-locked with GPU_Pin (:109-111).
    Inside every timed call: H2D of S and h (16 N^2 B), S^-1 h on the device (cuSOLVER potrf + potrs), 24 Lanczos passes
    for the spectral interval, the Chebyshev step of dt = 0.5 fs, D2H of H' (8 N^2 B, overlapped with the series), of
    the packets and of AO_bra.  Two propagators of the symbol are timed (DYNEMOL_B200_MODE):
      chebyshev25  the reference's chain of order-25 series (Chebyshev_gpu.cpp structure, rescaled) -> the e2e terms/s
      chebyshev    ONE expansion per step, order from the Bessel decay -> fewer passes for the same step (steps/s)."""
    from dynemol_b200 import api
    S = host_Sh[0].numpy().T; h = host_Sh[1].numpy().T              # symmetric: Fortran-ordered views of the pinned buffers
    Hp = np.zeros((N, N), dtype=np.float64, order="F")              # the caller's persistent H_prime ...
    api.gpu_pin(Hp)                                                 # ... pinned like ElHl_Chebyshev_GPU.f:111
    dt = 5e-4; tau_max = dt / H_BAR
    out = {"h2d_bytes_per_step": int(16 * N * N + 2 * 2 * 16 * N), "d2h_bytes_per_step": int(8 * N * N + 3 * 2 * 16 * N),
           "call": "propagationelhl2_gpucaller_ (pinned host S, h -> H', packets, AO_bra), dt=0.0005 ps", "modes": {}}
    try:
        for mode in ("chebyshev25", "chebyshev"):
            os.environ["DYNEMOL_B200_MODE"] = mode
            o = api.legacy_propagationelhl(S, h, Psi_bra, Psi_ket, 0.0, dt, tau_max, copy_inputs=False, out_H=Hp)   # first nuclear step (untimed)
            tau = np.minimum(tau_max, 1.15 * o["save_tau"])
            n_calls = max(1, args.e2e_steps)
            terms = 0
            t0 = time.perf_counter()
            for _ in range(n_calls):
                o = api.legacy_propagationelhl(S, h, Psi_bra, Psi_ket, 0.0, dt, tau, copy_inputs=False, out_H=Hp)
                terms += api.legacy_passes_last()
            t = time.perf_counter() - t0
            out["modes"][mode] = {"s_per_call": round(t / n_calls, 4), "terms_per_call": terms // n_calls, "terms_per_s": round(terms / t, 2),
                                  "nuclear_steps_per_s": round(n_calls / t, 4),
                                  "norm_el": float(abs(np.vdot(o["PSI_bra"][:, 0], o["PSI_ket"][:, 0])))}
    finally:
        os.environ.pop("DYNEMOL_B200_MODE", None)
        api.gpu_finalize()
        api.gpu_unpin(Hp)
    return out


def sharded_parity_check(dist, local_rank, dev, N=4096):
    """A short checked nuclear step on the shards of THIS run (Taylor, the order-25 Chebyshev chain and the single
    Chebyshev expansion) against the CPU oracle on rank 0: identical decision traces, wavepackets within 1e-10.
    The operator is formed once on rank 0 and broadcast, so every rank cuts its rows from the same bits."""
    import torch
    from dynemol_b200 import api, synthetic as syn
    from dynemol_b200.sharded import init_sharded
    rank, world = dist.get_rank(), dist.get_world_size()
    try:
        w = syn.make_workload(N)
        Hp_t = torch.empty((N, N), dtype=torch.float64, device=dev)
        if rank == 0:
            P1 = api.Propagator(N, device=local_rank)
            Hp = P1.form_hprime(w.S, w.h)
            P1.close()
            Hp_t.copy_(torch.from_numpy(np.ascontiguousarray(Hp)))
        dist.broadcast(Hp_t, src=0)
        Hp = np.asfortranarray(Hp_t.cpu().numpy())
        del Hp_t
        P, row0, m = init_sharded(N, dist, local_rank)
        P.upload_hprime(Hp)
        out = {"basis": N, "ok": True, "exchange": "p2p-fused" if P.info()["p2p"] else "nccl"}
        dt = 2e-7
        P.set_packets(w.Psi_bra, w.Psi_ket)
        lo, hi = P.estimate_spectral_bounds(24, 0.05)               # sharded Lanczos (collective)
        cases = [("taylor", api.MODE_TAYLOR, dt), ("chebyshev25", api.MODE_CHEBYSHEV, 50 * dt), ("chebyshev", api.MODE_CHEBYSHEV_FULL, 50 * dt)]
        for name, mode, dtc in cases:
            tau0 = dtc / api.H_BAR
            P.set_packets(w.Psi_bra, w.Psi_ket)
            save, traces = P.propagate(0.0, dtc, tau0, mode=mode)
            bra, ket = P.get_packets()
            if rank == 0:
                import oracle
                oracle.use_all_host_threads()
                worst = 0.0; same = True
                for p in range(2):
                    if mode == api.MODE_TAYLOR:
                        b, k, _, st, tr = oracle.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dtc, tau0)
                        same = same and [(e[0], e[1], e[2]) for e in traces[p].events()] == [(e[0], e[1], e[2]) for e in tr.events()] and save[p] == st
                    elif mode == api.MODE_CHEBYSHEV:
                        b, k, _, st, tr = oracle.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dtc, tau0, 0.5 * (hi + lo), 0.5 * (hi - lo))
                        same = same and [(e[0], e[1], e[2]) for e in traces[p].events()] == [(e[0], e[1], e[2]) for e in tr.events()] and save[p] == st
                    else:
                        from oracle import taylor_numpy as tn
                        b, k, n_terms, okn = tn.cheb_full_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dtc, 0.5 * (hi + lo), 0.5 * (hi - lo))
                        same = same and okn and traces[p].n_matvec_pairs == n_terms
                    worst = max(worst, np.abs(bra[:, p] - b).max() / np.abs(b).max(), np.abs(ket[:, p] - k).max() / np.abs(k).max())
                good = bool(same and worst < 1e-10)
                out[name] = {"ok": good, "same_decisions": bool(same), "worst_rel_err": float(worst), "terms": int(traces[0].n_matvec_pairs)}
                out["ok"] = bool(out["ok"] and good)
            dist.barrier()
        P.close()
        return out
    except Exception as e:              # the throughput line must survive a failure of the check -- but say so loudly
        return {"ok": False, "error": repr(e)[:300]}


def run_team_legacy(args):
    """Single-process multi-GPU through the PRIMARY symbol: propagationelhl2_gpucaller_ with DYNEMOL_B200_GPUS=P (the way a
    one-process Fortran caller reaches the GPUs of a box): pinned host S, h in; H', packets, AO_bra out; H'
    formed on the first GPU, row-sharded over P GPUs by peer copies, propagated with the fused NVLink exchange.
    Default size: BASELINE config 5 (N = 30720).  Run as a plain `python bench.py --team P` (NOT under torchrun)."""
    import torch
    from dynemol_b200 import api, synthetic as syn
    N, P = args.basis or 30720, args.team
    dev = torch.device("cuda", 0)
    t0 = time.time()
    S_t, h_t, _ = syn.make_S_h_torch(N, dev)
    Sh = torch.empty((N, N), dtype=torch.float64, pin_memory=True); hh = torch.empty((N, N), dtype=torch.float64, pin_memory=True)
    Sh.copy_(S_t); hh.copy_(h_t)
    w = 64
    C = torch.zeros((N, 2), dtype=torch.float64, device=dev)
    C[0:w, 0] = torch.tensor(np.random.default_rng(42).normal(size=w), device=dev)
    C[w:2 * w, 1] = torch.tensor(np.random.default_rng(43).normal(size=w), device=dev)
    SC = S_t @ C
    nrm = torch.sqrt((C * SC).sum(0))
    Psi_ket = np.asfortranarray((C / nrm).cpu().numpy().astype(np.complex128))
    Psi_bra = np.asfortranarray((SC / nrm).cpu().numpy().astype(np.complex128))
    del S_t, h_t, SC, C
    torch.cuda.empty_cache()
    gen_s = time.time() - t0
    S = Sh.numpy().T; h = hh.numpy().T
    Hp = np.zeros((N, N), dtype=np.float64, order="F")
    api.gpu_pin(Hp)
    dt = 5e-4; tau_max = dt / H_BAR
    res = {}
    os.environ["DYNEMOL_B200_GPUS"] = str(P)
    try:
        for mode in ("chebyshev", "chebyshev25"):
            os.environ["DYNEMOL_B200_MODE"] = mode
            t1 = time.perf_counter()
            o = api.legacy_propagationelhl(S, h, Psi_bra, Psi_ket, 0.0, dt, tau_max, copy_inputs=False, out_H=Hp)   # first step (+ team set-up)
            first_s = time.perf_counter() - t1
            tau = np.minimum(tau_max, 1.15 * o["save_tau"])
            n_calls = max(1, args.e2e_steps)
            terms = 0
            t1 = time.perf_counter()
            for _ in range(n_calls):
                o = api.legacy_propagationelhl(S, h, Psi_bra, Psi_ket, 0.0, dt, tau, copy_inputs=False, out_H=Hp)
                terms += api.legacy_passes_last()
            t = time.perf_counter() - t1
            resid = float(np.abs(S[:, :512] @ Hp[:, :8] - h[:, :8]).max() / np.abs(h[:, :8]).max()) if N <= 512 else \
                float(np.abs(S @ Hp[:, :4] - h[:, :4]).max() / np.abs(h[:, :4]).max())
            res[mode] = {"s_per_call": round(t / n_calls, 4), "terms_per_call": terms // n_calls, "terms_per_s": round(terms / t, 1),
                         "nuclear_steps_per_s": round(n_calls / t, 4), "first_call_s": round(first_s, 2),
                         "norm_el": float(abs(np.vdot(o["PSI_bra"][:, 0], o["PSI_ket"][:, 0]))), "S_Hprime_minus_h_rel": resid}
    finally:
        os.environ.pop("DYNEMOL_B200_MODE", None); os.environ.pop("DYNEMOL_B200_GPUS", None)
        api.gpu_finalize()
        api.gpu_unpin(Hp)
    line = {"metric": "nuclear step (dt = 0.5 fs) through propagationelhl2_gpucaller_ on a single-process team", "unit": "s per call", "n_gpus": P,
            "config": {"workload": "synthetic EHT Hamiltonian N=%d basis, host S, h -> H', packets, AO_bra; DYNEMOL_B200_GPUS=%d" % (N, P), "basis": N,
                       "h2d_bytes_per_call": int(16 * N * N + 2 * 2 * 16 * N), "d2h_bytes_per_call": int(8 * N * N + 3 * 2 * 16 * N), "gen_s": round(gen_s, 1)},
            "modes": res, "dtype": "f64", "data": "synthetic"}
    print(json.dumps(line))


def cpu_baseline(Hp_np, Psi_bra, Psi_ket, tau, budget_s=15.0):
    import oracle                                                    # the checker, timed as the CPU baseline only
    oracle.build()
    oracle.use_all_host_threads()
    N = Hp_np.shape[0]
    t0 = time.perf_counter(); oracle.terms(Hp_np, Psi_bra, Psi_ket, tau, 2); t2 = time.perf_counter() - t0
    n = int(max(2, min(200, budget_s / max(t2 / 2, 1e-6))))
    t0 = time.perf_counter(); oracle.terms(Hp_np, Psi_bra, Psi_ket, tau, n); t = time.perf_counter() - t0
    return {"value": round(n / t, 3), "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d el+hole terms (4 dzgemv each, Taylor.f:94-95 x 2 particles) at N=%d, g++ -O3 OpenMP, %.1f s" % (n, N, t),
            "gbs_reference_style": round(4 * 8.0 * N * N * n / t / 1e9, 1)}


def host_mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def run_reference(args):
    """The reference's own CPU algorithm for the path (oracle port: the Fortran/MKL original cannot be built here),
    all host threads, on the same config/metric; each step a bounded sample."""
    import oracle
    oracle.build()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle.use_all_host_threads()            # torchrun exports OMP_NUM_THREADS=1 to its workers; the other ranks are idle here
    N = args.basis or (16384 if args.gpus == 1 else 65536)
    # the workload's own size when the host can hold it (8 N^2 B + slack), else N = 16384 scaled by the byte ratio
    N_run = N if host_mem_available_gb() > 8.0 * N * N / 1e9 * 1.3 + 8.0 else min(N, 16384)
    rng = np.random.default_rng(1)
    # CPU GEMV time does not depend on the matrix values: a random block, tiled, stands in for H' (filling 34 GB with
    # fresh random numbers would take longer than the measurement)
    blk = rng.standard_normal((N_run, 256)) * 1e-2
    Hp = np.empty((N_run, N_run), dtype=np.float64, order="F")
    for j in range(0, N_run, 256):
        Hp[:, j:j + 256] = blk[:, :min(256, N_run - j)]
    Psi = np.asfortranarray(rng.standard_normal((N_run, 2)) + 1j * rng.standard_normal((N_run, 2)))
    tau = 1e-4
    per_step = max(1, args.ref_terms_per_step)
    # bound the whole run: at most ~150 s of timed CPU work whatever --steps says (each step is a sample of the workload)
    t0 = time.perf_counter(); oracle.terms(Hp, Psi, Psi, tau, 1); t1 = time.perf_counter() - t0
    warm = max(0, args.warmup - 1)
    for _ in range(min(warm, max(0, int(20.0 / max(t1 * per_step, 1e-6))))):
        oracle.terms(Hp, Psi, Psi, tau, per_step)
    steps = max(1, min(args.steps, int(150.0 / max(t1 * per_step, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.terms(Hp, Psi, Psi, tau, per_step)
    t = time.perf_counter() - t0
    rate = steps * per_step / t
    scale = (N_run / N) ** 2                 # bytes per term scale with N^2 (memory-bound GEMV); 1 when measured at full size
    value = rate * scale
    cores = oracle.num_threads()
    sample = "%d steps x %d el+hole terms at N=%d%s, oracle port (g++ -O3 OpenMP; MKL is not available), %d threads" % (
        steps, per_step, N_run, "" if N_run == N else " (scaled by (N_run/N)^2 to N=%d: host RAM too small for 8 N^2 B)" % N, cores)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1 + warm, "ms_per_step": round(1e3 * t / steps, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N, args.gpus), "basis": N, "terms_per_step": per_step,
                       "measured_at_basis": N_run, "impl": "CPU oracle port, OpenMP, %d threads" % cores},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(N, gpus):
    """One workload string for both arms (the driver compares `config.workload` of the two lines)."""
    if gpus == 1:
        return "synthetic EHT Hamiltonian N=%d basis, el+hole packets, Chebyshev dt=0.5 fs" % N
    return "synthetic EHT Hamiltonian N=%d basis, row-sharded H' (el+hole packets)" % N


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--basis", type=int, default=0)
    ap.add_argument("--e2e-dt", type=float, default=0.0, help="nuclear step (ps) of the end-to-end call (default 5e-4 Chebyshev / 5e-6 Taylor)")
    ap.add_argument("--e2e-mode", default="cheb", choices=["cheb", "taylor"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--ref-terms-per-step", type=int, default=2)
    ap.add_argument("--kernel", default="tma", choices=["tma", "ldg"])
    ap.add_argument("--no-ref1", action="store_true", help="multi-GPU: skip the 1-GPU same-workload reference on rank 0")
    ap.add_argument("--no-parity", action="store_true", help="multi-GPU: skip the checked N=4096 step on the shards")
    ap.add_argument("--config5", action="store_true", help="multi-GPU: BASELINE config 5 (N~30k, Taylor vs Chebyshev nuclear step) instead of the throughput line")
    ap.add_argument("--config5-taylor-frac", type=float, default=0.02, help="fraction of the 0.5 fs step the Taylor propagator is timed on (scaled linearly)")
    ap.add_argument("--team", type=int, default=0, help="single process: time the primary legacy symbol on a team of this many GPUs (DYNEMOL_B200_GPUS)")
    ap.add_argument("--skip-small", action="store_true", help="N=1: skip the N=900 small-operator side measurement")
    ap.add_argument("--skip-65k", action="store_true", help="N=1: skip the N=65536 single-GPU side measurement")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: no CPU baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: no end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.team > 1:
        return run_team_legacy(args)
    if args.gpus == 1:
        return run_ours_single(args)
    from dynemol_b200 import sharded
    if args.config5:
        return sharded.config5_main(args)
    return sharded.bench_main(args)


if __name__ == "__main__":
    main()
