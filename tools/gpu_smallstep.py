#!/usr/bin/env python
"""Wall time of whole nuclear steps (dyb_propagate) for a small operator: how much is series time, how much host overhead."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from dynemol_b200 import api, synthetic as syn
H_BAR = 6.58264e-4
N = int(sys.argv[1]) if len(sys.argv) > 1 else 900
w = syn.make_workload(N)
P = api.Propagator(N)
P.form_hprime(w.S, w.h, want_hprime=False)
for mode, dt in ((api.MODE_TAYLOR, 2e-5), (api.MODE_CHEBYSHEV, 5e-4)):
    for kind in ("term", "auto"):
        P.set_series_kernel(kind)
        P.set_packets(w.Psi_bra, w.Psi_ket)
        if mode == api.MODE_CHEBYSHEV:
            P.estimate_spectral_bounds(24, 0.05)
        tau = dt / H_BAR
        save, tr = P.propagate(0.0, dt, tau, mode=mode)          # first step finds tau
        l0 = P.launch_count()
        t0 = time.perf_counter()
        steps = 20
        for s in range(steps):
            save, tr = P.propagate(0.0, dt, np.minimum(tau, 1.15 * save), mode=mode)
        P.sync()
        t = (time.perf_counter() - t0) / steps
        terms = P.info()["passes_last"]
        n_series = tr[0].n_convergence_calls + sum(1 for e in tr[0].events() if e[0] == 2)
        print(f"N={N} mode={'taylor' if mode == api.MODE_TAYLOR else 'cheb'} dt={dt} series={kind}: {t*1e3:.3f} ms/step, {terms} passes/step, "
              f"{t*1e6/max(terms,1):.2f} us per pass all-in, launches/step {(P.launch_count()-l0)/steps:.0f}, el series ~{n_series}")
P.close()
