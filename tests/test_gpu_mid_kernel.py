"""GPU suite: the streamed one-launch series kernel for mid-size operators (DYB_SERIES_MID, csrc/mid.cuh).

It must be interchangeable with the two-launch path (DYB_SERIES_PER_TERM), which the other GPU tests pin against the oracle
at every size: identical decision traces (tau schedule, exit index of every Convergence call of Taylor.f:132-219, sub-step
count of Taylor.f:81-126), wavepackets within 1e-10 of the oracle (the north_star tolerance) and within 1e-12 of the
two-launch path (only the summation order differs).  Sizes: its natural range (1828 = the first size (4 orbitals per atom) past the resident
kernel, 2304, 3000, 4096, 5000, 6144 = the last one, with the 16-bit collect table: different grid shapes, ragged last block
row / column) and small operators forced through it (300, 700: few block rows, many block columns, whole tiles out of
bounds)."""
import os

import numpy as np
import pytest

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
H_BAR = 6.58264e-4
SERIES_PER_TERM, SERIES_RESIDENT, SERIES_MID = 1, 3, 5


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import api as a
    assert a.device_count() > 0
    return a


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def events3(tr):
    return [(e[0], e[1], e[2]) for e in tr.events()]


def run(api, kind, Hp, bra, ket, dt, tau0, mode=None, bounds=None):
    N = Hp.shape[0]
    P = api.Propagator(N)
    P.set_series_kernel(kind)
    P.upload_hprime(Hp)
    P.set_packets(bra, ket)
    if bounds is not None:
        P.set_spectral_bounds(*bounds)
    in_effect = P.info()["series_kernel"]
    l0 = P.launch_count()
    save_tau, traces = P.propagate(0.0, dt, tau0, mode=api.MODE_TAYLOR if mode is None else mode)
    n_launch = P.launch_count() - l0
    b, k = P.get_packets()
    P.close()
    return in_effect, save_tau, traces, b, k, n_launch


def hprime(oracle_mod, w):
    """H' = S^-1 h: the oracle's invert + multiply for small operators, the library's own formation (pinned against the
    oracle in test_gpu_parity.py) for the large ones -- both runs and the oracle then use the same matrix."""
    N = w.S.shape[0]
    if N <= 1024:
        return oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    from dynemol_b200 import api
    P = api.Propagator(N)
    Hp = P.form_hprime(w.S, w.h)
    P.close()
    return Hp


def test_mid_is_the_default_between_resident_and_streaming(api):
    for N, want in ((900, SERIES_RESIDENT), (1825, SERIES_MID), (2304, SERIES_MID), (4096, SERIES_MID), (6144, SERIES_MID), (6400, SERIES_PER_TERM), (16384, SERIES_PER_TERM)):
        P = api.Propagator(N)
        assert P.info()["series_kernel"] == want, (N, P.info())
        P.close()


@pytest.mark.parametrize("N,dt", [(300, 1e-5), (700, 2e-6), (1828, 6e-7), (2304, 5e-7), (3000, 4e-7), (4096, 2e-7), (5000, 2e-7), (6144, 1.5e-7)])
def test_mid_taylor_matches_per_term_and_oracle(api, oracle_mod, N, dt):
    w = syn.make_workload(N)
    Hp = hprime(oracle_mod, w)
    tau0 = dt / H_BAR
    kind_m, st_m, tr_m, b_m, k_m, nl_m = run(api, "mid", Hp, w.Psi_bra, w.Psi_ket, dt, tau0)
    kind_t, st_t, tr_t, b_t, k_t, nl_t = run(api, "term", Hp, w.Psi_bra, w.Psi_ket, dt, tau0)
    assert kind_m == SERIES_MID and kind_t == SERIES_PER_TERM
    assert nl_m < nl_t
    for p in range(2):
        assert events3(tr_m[p]) == events3(tr_t[p])
        assert st_m[p] == st_t[p]
        assert tr_m[p].n_matvec_pairs == tr_t[p].n_matvec_pairs
        assert relerr(b_m[:, p], b_t[:, p]) < 1e-12 and relerr(k_m[:, p], k_t[:, p]) < 1e-12
        b, k, _, st, tr = oracle_mod.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        assert events3(tr_m[p]) == events3(tr) and st_m[p] == st
        assert relerr(b_m[:, p], b) < 1e-10 and relerr(k_m[:, p], k) < 1e-10


@pytest.mark.parametrize("N", [700, 2304])
def test_mid_chebyshev_matches_per_term_and_oracle(api, oracle_mod, N):
    """Three-term recurrence + spectral rescaling through the mid kernel (dt = 0.5 fs, BASELINE config 3's step): the
    order-25 chain with its steady sub-steps chained inside one launch."""
    dt = 5e-4
    w = syn.make_workload(N)
    Hp = hprime(oracle_mod, w)
    e = np.linalg.eigvals(Hp).real
    pad = 0.02 * (e.max() - e.min())
    bounds = (e.min() - pad, e.max() + pad)
    tau0 = dt / H_BAR
    (km, st_m, tr_m, b_m, k_m, _) = run(api, "mid", Hp, w.Psi_bra, w.Psi_ket, dt, tau0, mode=api.MODE_CHEBYSHEV, bounds=bounds)
    (kt, st_t, tr_t, b_t, k_t, _) = run(api, "term", Hp, w.Psi_bra, w.Psi_ket, dt, tau0, mode=api.MODE_CHEBYSHEV, bounds=bounds)
    assert km == SERIES_MID and kt == SERIES_PER_TERM
    ebar, de = 0.5 * (bounds[0] + bounds[1]), 0.5 * (bounds[1] - bounds[0])
    for p in range(2):
        assert events3(tr_m[p]) == events3(tr_t[p]) and st_m[p] == st_t[p]
        assert relerr(b_m[:, p], b_t[:, p]) < 1e-12 and relerr(k_m[:, p], k_t[:, p]) < 1e-12
        if N <= 1000:
            b, k, _, st, tr = oracle_mod.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0, ebar, de)
            assert events3(tr_m[p]) == events3(tr)
            assert relerr(b_m[:, p], b) < 1e-10 and relerr(k_m[:, p], k) < 1e-10


def test_mid_single_expansion_matches_expm(api, oracle_mod):
    """DYB_MODE_CHEBYSHEV_FULL (one expansion of ~R terms per step) in ONE launch of the mid kernel, against expm."""
    from scipy.linalg import expm
    N, dt = 600, 5e-4
    w = syn.make_workload(N)
    Hp = hprime(oracle_mod, w)
    P = api.Propagator(N)
    P.set_series_kernel("mid")
    assert P.info()["series_kernel"] == SERIES_MID
    P.upload_hprime(Hp)
    P.set_packets(w.Psi_bra, w.Psi_ket)
    P.estimate_spectral_bounds(n_iter=24, margin=0.05)
    tau0 = dt / H_BAR
    l0 = P.launch_count()
    P.propagate(0.0, dt, tau0, mode=api.MODE_CHEBYSHEV_FULL)
    assert P.launch_count() - l0 <= 6
    bra, ket = P.get_packets()
    U = expm(-1j * tau0 * Hp)
    for p in range(2):
        assert relerr(ket[:, p], U @ w.Psi_ket[:, p]) < 1e-10
        assert relerr(bra[:, p], U.T @ w.Psi_bra[:, p]) < 1e-10
    P.close()


def test_mid_single_particle_and_unequal_series(api, oracle_mod):
    """Electron alone, and electron + hole with different time steps: independent state machines inside the shared launch."""
    N, dt = 640, 4e-6
    w = syn.make_workload(N)
    Hp = hprime(oracle_mod, w)
    P = api.Propagator(N)
    P.set_series_kernel("mid")
    P.upload_hprime(Hp)
    assert P.info()["series_kernel"] == SERIES_MID
    P.set_packets(w.Psi_bra[:, 0], w.Psi_ket[:, 0])
    tau0 = dt / H_BAR
    st1, tr1 = P.propagate(0.0, dt, tau0)
    b1, k1 = P.get_packets()
    b, k, _, st, tr = oracle_mod.propagation(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], 0.0, dt, tau0)
    assert events3(tr1[0]) == events3(tr) and relerr(b1[:, 0], b) < 1e-10 and relerr(k1[:, 0], k) < 1e-10
    P.set_packets(w.Psi_bra, w.Psi_ket)
    taus = np.array([tau0, 0.37 * tau0])
    st2, tr2 = P.propagate(0.0, dt, taus)
    b2, k2 = P.get_packets()
    for p in range(2):
        b, k, _, st, tr = oracle_mod.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, float(taus[p]))
        assert events3(tr2[p]) == events3(tr) and st2[p] == st
        assert relerr(b2[:, p], b) < 1e-10 and relerr(k2[:, p], k) < 1e-10
    assert np.array_equal(b2[:, 0], b1[:, 0]) and np.array_equal(k2[:, 0], k1[:, 0]), "the hole must not perturb the electron"
    P.close()


def test_mid_chained_steady_loop_is_bit_identical(api):
    """The whole steady loop of a nuclear step in ONE launch of the mid kernel against the same kernel driven one sub-step
    per launch (DYNEMOL_B200_CHAIN=0): identical tau schedule, events, pass counts and wavepackets, bit for bit, over several
    nuclear steps with the carried-over tau (ElHl_Chebyshev.f:182-184; provokes failed norm tests and rescaling)."""
    N = 600
    w = syn.make_workload(N)
    res = {}
    bounds = None
    for chain in ("0", "1"):
        os.environ["DYNEMOL_B200_CHAIN"] = chain
        try:
            P = api.Propagator(N)
        finally:
            os.environ.pop("DYNEMOL_B200_CHAIN", None)
        P.set_series_kernel("mid")
        assert P.info()["series_kernel"] == SERIES_MID
        P.form_hprime(w.S, w.h, want_hprime=False)
        out = []
        for mode, dt in ((api.MODE_TAYLOR, 2e-5), (api.MODE_CHEBYSHEV, 5e-4)):
            P.set_packets(w.Psi_bra, w.Psi_ket)
            if mode == api.MODE_CHEBYSHEV:
                if bounds is None:
                    bounds = P.estimate_spectral_bounds(24, 0.05)
                P.set_spectral_bounds(*bounds)
            tau = dt / H_BAR
            save = np.array([tau, tau])
            for _ in range(3):
                l0 = P.launch_count()
                save, tr = P.propagate(0.0, dt, np.minimum(tau, 1.15 * save), mode=mode)
                b, k = P.get_packets()
                out.append(dict(save=save.copy(), events=[[tuple(e) for e in t.events()] for t in tr], pairs=[t.n_matvec_pairs for t in tr],
                                sub=[t.n_substeps for t in tr], resc=[t.n_rescale for t in tr], bra=b.copy(), ket=k.copy(), launches=P.launch_count() - l0))
        res[chain] = out
        P.close()
    n_fail = 0
    for i, (a, b) in enumerate(zip(res["0"], res["1"])):
        assert np.array_equal(a["save"], b["save"])
        assert a["events"] == b["events"] and a["pairs"] == b["pairs"] and a["sub"] == b["sub"] and a["resc"] == b["resc"]
        assert np.array_equal(a["bra"], b["bra"]) and np.array_equal(a["ket"], b["ket"])
        assert b["launches"] < a["launches"]
        n_fail += sum(a["resc"])
        assert max(a["sub"]) >= 10, "the case must actually have a steady loop worth chaining"
    assert n_fail > 0, "the case must exercise a failed sub-step inside a chained launch"


def test_mid_is_bitwise_repeatable(api, oracle_mod):
    """No atomics, fixed summation orders: two runs of the same step give the same bits."""
    N, dt = 2304, 5e-7
    w = syn.make_workload(N)
    Hp = hprime(oracle_mod, w)
    a = run(api, "mid", Hp, w.Psi_bra, w.Psi_ket, dt, dt / H_BAR)
    b = run(api, "mid", Hp, w.Psi_bra, w.Psi_ket, dt, dt / H_BAR)
    assert a[0] == SERIES_MID
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]) and np.array_equal(a[1], b[1])
