"""GPU suite: the Chebyshev mode (DYB_MODE_CHEBYSHEV) = the reference's un-linked Chebyshev series
(Chebyshev_gpu.cpp:347-485,524-643) on the spectrally rescaled operator, against the oracle's statement of
the same algorithm (golden fixtures cross-checked with numpy and expm) and against expm directly."""
import os

import numpy as np
import pytest
from scipy.linalg import expm

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
H_BAR = 6.58264e-4


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import api as a
    assert a.device_count() > 0
    return a


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", ["cheb_N64_dt5e-4", "cheb_N128_dt5e-5"])
def test_chebyshev_matches_golden(api, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    N = int(g["N"])
    P = api.Propagator(N)
    P.upload_hprime(g["H_prime"])
    P.set_packets(g["Psi_bra0"], g["Psi_ket0"])
    P.set_spectral_bounds(float(g["emin"]), float(g["emax"]))
    save_tau, traces = P.propagate(float(g["t_init"]), float(g["t_max"]), float(g["tau0"]), mode=api.MODE_CHEBYSHEV)
    bra, ket = P.get_packets()
    for p, tag in enumerate(("el", "hl")):
        ev = np.array([[e[0], e[1], e[2]] for e in traces[p].events()])
        assert np.array_equal(ev, g[f"{tag}_events"]), "decision trace differs from the oracle's"
        assert traces[p].n_matvec_pairs == int(g[f"{tag}_matvec_pairs"])
        assert abs(save_tau[p] - float(g[f"{tag}_save_tau"])) <= 1e-15 * save_tau[p]
        assert relerr(bra[:, p], g[f"{tag}_bra"]) < 1e-10
        assert relerr(ket[:, p], g[f"{tag}_ket"]) < 1e-10
    P.close()


def test_lanczos_bounds_and_half_femtosecond_step(api, oracle_mod):
    """dt = 0.5 fs (BASELINE config 3's step) with bounds estimated on the device: the result must agree with the
    exact propagator to the accuracy the 1e-8 tolerances give, using ~25x fewer dual products than the Taylor mode."""
    N, dt = 512, 5e-4
    w = syn.make_workload(N)
    Hp = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    e = np.linalg.eigvals(Hp).real
    P = api.Propagator(N)
    P.upload_hprime(Hp)
    P.set_packets(w.Psi_bra, w.Psi_ket)
    lo, hi = P.estimate_spectral_bounds(n_iter=60, margin=0.02)
    assert (lo, hi) == P.estimate_spectral_bounds(n_iter=60, margin=0.02), "the estimate must be bit-reproducible"
    width = e.max() - e.min()
    assert lo <= e.min() + 1e-6 * width and hi >= e.max() - 1e-6 * width, (lo, hi, e.min(), e.max())
    assert hi - lo < 1.10 * width, "bounds should be tight, not Gershgorin-loose"
    tau0 = dt / H_BAR
    save_tau, traces = P.propagate(0.0, dt, tau0, mode=api.MODE_CHEBYSHEV)
    bra, ket = P.get_packets()
    U = expm(-1j * tau0 * Hp)
    for p in range(2):
        assert np.abs(U @ w.Psi_ket[:, p] - ket[:, p]).max() < 2e-7
        assert np.abs(U.T @ w.Psi_bra[:, p] - bra[:, p]).max() < 2e-7
        assert abs(abs(np.vdot(bra[:, p], ket[:, p])) - 1.0) < 1e-7
    cheb_terms = P.info()["passes_last"]
    # same bounds through the oracle's statement of the algorithm: identical decisions
    ebar, de = 0.5 * (hi + lo), 0.5 * (hi - lo)
    for p in range(2):
        b, k, _, st, tr = oracle_mod.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0, ebar, de)
        assert [(x[0], x[1], x[2]) for x in traces[p].events()] == [(x[0], x[1], x[2]) for x in tr.events()]
        assert relerr(bra[:, p], b) < 1e-10 and relerr(ket[:, p], k) < 1e-10
    assert cheb_terms < 4000, cheb_terms
    P.close()
