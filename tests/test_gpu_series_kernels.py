"""GPU suite: the three ways a series can be launched (include/dynemol_b200.h, DYB_SERIES_*) must be
interchangeable.  PER_TERM (two launches per term) is the path the other GPU tests pin against the oracle at every
size; here the single-launch kernels -- RESIDENT (H' blocked over the shared memories, small N, csrc/resident.cuh)
and STREAM (cooperative TMA kernel, csrc/series.cuh) -- are compared with it and with the oracle directly:
identical decision traces (tau schedule, exit index of every Convergence call, sub-step count), wavepackets within
1e-10 of the oracle and within 1e-12 of each other (only the summation order differs)."""
import numpy as np
import pytest

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
H_BAR = 6.58264e-4
SERIES_PER_TERM, SERIES_STREAM, SERIES_RESIDENT = 1, 2, 3


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
    save_tau, traces = P.propagate(0.0, dt, tau0, mode=api.MODE_TAYLOR if mode is None else mode)
    b, k = P.get_packets()
    n_launch = P.launch_count() if hasattr(P, "launch_count") else None
    P.close()
    return in_effect, save_tau, traces, b, k, n_launch


# N = 36: a single CTA; 256: 8 x 8 CTAs of 32-row blocks; 892: about the heptazine-in-water QM region,
# not a multiple of the grid side; 1824: the largest operator that fits (152-row blocks on a 12 x 12 grid)
@pytest.mark.parametrize("N,dt", [(36, 2e-5), (256, 5e-6), (892, 2e-6), (1824, 5e-7)])
def test_resident_taylor_matches_per_term_and_oracle(api, oracle_mod, N, dt):
    w = syn.make_workload(N)
    Hp = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    tau0 = dt / H_BAR
    kind_r, st_r, tr_r, b_r, k_r, _ = run(api, "resident", Hp, w.Psi_bra, w.Psi_ket, dt, tau0)
    kind_t, st_t, tr_t, b_t, k_t, _ = run(api, "term", Hp, w.Psi_bra, w.Psi_ket, dt, tau0)
    assert kind_r == SERIES_RESIDENT and kind_t == SERIES_PER_TERM
    for p in range(2):
        assert events3(tr_r[p]) == events3(tr_t[p])
        assert st_r[p] == st_t[p]
        assert tr_r[p].n_matvec_pairs == tr_t[p].n_matvec_pairs
        assert relerr(b_r[:, p], b_t[:, p]) < 1e-12 and relerr(k_r[:, p], k_t[:, p]) < 1e-12
        b, k, _, st, tr = oracle_mod.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        assert events3(tr_r[p]) == events3(tr) and st_r[p] == st
        assert relerr(b_r[:, p], b) < 1e-10 and relerr(k_r[:, p], k) < 1e-10


def test_resident_is_the_default_for_small_operators(api):
    P = api.Propagator(900)
    info = P.info()
    assert info["series_kernel"] == SERIES_RESIDENT
    assert info["resident_grid_side"] == 12 and info["resident_block"] == 75
    P.close()
    P = api.Propagator(4096)                     # does not fit: 342-row blocks -> the streamed one-launch kernel (mid.cuh)
    assert P.info()["series_kernel"] == 5
    P.set_series_kernel("resident")              # asking for it does not force it
    assert P.info()["series_kernel"] == SERIES_PER_TERM
    P.close()


def test_resident_chebyshev_matches_per_term(api, oracle_mod):
    """Three-term recurrence + spectral rescaling through the resident kernel (dt = 0.5 fs, BASELINE config 3's step)."""
    N, dt = 700, 5e-4
    w = syn.make_workload(N)
    Hp = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    e = np.linalg.eigvals(Hp).real
    pad = 0.02 * (e.max() - e.min())
    bounds = (e.min() - pad, e.max() + pad)
    tau0 = dt / H_BAR
    out = {}
    for kind in ("resident", "term"):
        out[kind] = run(api, kind, Hp, w.Psi_bra, w.Psi_ket, dt, tau0, mode=api.MODE_CHEBYSHEV, bounds=bounds)
    (kr, st_r, tr_r, b_r, k_r, _), (kt, st_t, tr_t, b_t, k_t, _) = out["resident"], out["term"]
    assert kr == SERIES_RESIDENT and kt == SERIES_PER_TERM
    ebar, de = 0.5 * (bounds[0] + bounds[1]), 0.5 * (bounds[1] - bounds[0])
    for p in range(2):
        assert events3(tr_r[p]) == events3(tr_t[p]) and st_r[p] == st_t[p]
        assert relerr(b_r[:, p], b_t[:, p]) < 1e-12 and relerr(k_r[:, p], k_t[:, p]) < 1e-12
        b, k, _, st, tr = oracle_mod.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0, ebar, de)
        assert events3(tr_r[p]) == events3(tr)
        assert relerr(b_r[:, p], b) < 1e-10 and relerr(k_r[:, p], k) < 1e-10


def test_resident_single_particle_and_unequal_series(api, oracle_mod):
    """Electron alone, and electron + hole with different time steps (different series lengths, one particle
    finishing earlier): the particles stay independent state machines inside the shared launch."""
    N, dt = 300, 1e-5
    w = syn.make_workload(N)
    Hp = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    P = api.Propagator(N)
    P.upload_hprime(Hp)
    assert P.info()["series_kernel"] == SERIES_RESIDENT
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


def test_chained_steady_loop_is_bit_identical(api):
    """Small operators run the whole steady loop of a nuclear step (Taylor.f:81-126) in ONE launch: the host predicts the
    sub-step schedule, the device chains the sub-steps (PartPass::begin / chain) and stops at a failed norm test.  Against
    the same kernel driven one sub-step per launch (DYNEMOL_B200_CHAIN=0): identical tau schedule, events, pass counts
    and wavepackets, bit for bit, over several nuclear steps with the carried-over tau of ElHl_Chebyshev.f:182-184 (which
    provokes failed norm tests and rescaling), in both modes; and far fewer launches."""
    import os
    N = 600
    w = syn.make_workload(N)
    res = {}
    bounds = None          # estimated once: the Lanczos estimate (host OpenMP reductions) is not bit-reproducible run to run
    for chain in ("0", "1"):
        os.environ["DYNEMOL_B200_CHAIN"] = chain
        try:
            P = api.Propagator(N)
        finally:
            os.environ.pop("DYNEMOL_B200_CHAIN", None)
        assert P.info()["series_kernel"] == SERIES_RESIDENT
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
        if i < 3:                                                # Taylor steps: hundreds of short sub-steps, a handful of launches
            assert b["launches"] < a["launches"] / 3
        n_fail += sum(a["resc"])
        assert max(a["sub"]) >= 10, "the case must actually have a steady loop worth chaining"
    assert n_fail > 0, "the case must exercise a failed sub-step inside a chained launch"
