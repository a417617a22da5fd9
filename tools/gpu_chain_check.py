#!/usr/bin/env python
"""Chained steady loop (one launch) against the unchained resident path: traces and packets must be bit-identical."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from dynemol_b200 import api, synthetic as syn
H_BAR = 6.58264e-4
N = int(sys.argv[1]) if len(sys.argv) > 1 else 900
w = syn.make_workload(N)
res = {}
for tag, chain in (("0", "0"), ("1", "1"), ("1b", "1"), ("0b", "0")):
    os.environ["DYNEMOL_B200_CHAIN"] = chain
    P = api.Propagator(N)
    P.form_hprime(w.S, w.h, want_hprime=False)
    out = []
    for mode, dt in ((api.MODE_TAYLOR, 2e-5), (api.MODE_CHEBYSHEV, 5e-4)):
        P.set_packets(w.Psi_bra, w.Psi_ket)
        if mode == api.MODE_CHEBYSHEV:
            print("bounds", P.estimate_spectral_bounds(24, 0.05))
        tau = dt / H_BAR
        save = np.array([tau, tau])
        for s in range(4):
            save, tr = P.propagate(0.0, dt, np.minimum(tau, 1.15 * save), mode=mode)
            b, k = P.get_packets()
            out.append((save.copy(), [[(e[0], e[1], e[2], e[3]) for e in t.events()] for t in tr], [t.n_matvec_pairs for t in tr], b.copy(), k.copy(), P.info()["passes_last"]))
    res[tag] = out
    P.close()
def maxdiff(x, y):
    d = np.abs(x - y); i = np.unravel_index(np.argmax(d), d.shape); return float(d.max()), i
for ta, tb in (("0", "0b"), ("1", "1b"), ("0", "1")):
    for i, (a, b) in enumerate(zip(res[ta], res[tb])):
        print(ta, tb, i, "bra", np.array_equal(a[3], b[3]), maxdiff(a[3], b[3]), "ket", np.array_equal(a[4], b[4]), maxdiff(a[4], b[4]), "events", a[1] == b[1])
for i, (a, b) in enumerate(zip(res["0"], res["1"])):
    same_ev = a[1] == b[1]
    print(i, "save", np.array_equal(a[0], b[0]), "events", same_ev, "pairs", a[2], b[2], "bra", np.array_equal(a[3], b[3]), "ket", np.array_equal(a[4], b[4]), "passes", a[5], b[5],
          "n_events", [len(x) for x in a[1]], [len(x) for x in b[1]])
    if not same_ev:
        for p in range(2):
            for j, (x, y) in enumerate(zip(a[1][p], b[1][p])):
                if x != y:
                    print("  first diff particle", p, "event", j, x, y); break
