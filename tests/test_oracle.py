"""CPU suite: pins the oracle (oracle/elhl_oracle.cpp) against everything available:
committed golden fixtures, an independent numpy transcription, scipy expm, and the parts of the
real reference that compile here (oracle/_ref).  No GPU needed."""
import os

import numpy as np
import pytest
from scipy.linalg import expm

from dynemol_b200 import synthetic as syn
from oracle import taylor_numpy as tn


def _hprime(oracle_mod, w):
    return oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)


def test_coefficients_follow_taylor_f(oracle_mod):
    tau = 0.37
    c = oracle_mod.coefficient(tau)
    k = np.arange(25)
    from math import factorial
    expect = np.array([(-1j * tau) ** int(i) / factorial(int(i)) for i in k])
    assert np.allclose(c, expect, rtol=1e-14, atol=0)
    assert np.allclose(c, tn.coefficient(tau), rtol=1e-15, atol=0)


@pytest.mark.parametrize("name", ["prop_N64_dt5e-6", "prop_N64_dt5e-4", "prop_N128_dt2e-5"])
def test_oracle_reproduces_golden(oracle_mod, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    Hp = g["H_prime"]
    for p, tag in enumerate(("el", "hl")):
        b, k, tau_out, save_tau, tr = oracle_mod.propagation(Hp, g["Psi_bra0"][:, p], g["Psi_ket0"][:, p],
                                                             float(g["t_init"]), float(g["t_max"]), float(g["tau0"]))
        assert np.abs(b - g[f"{tag}_bra"]).max() < 1e-14
        assert np.abs(k - g[f"{tag}_ket"]).max() < 1e-14
        assert tr.n_matvec_pairs == int(g[f"{tag}_matvec_pairs"])
        assert save_tau == float(g[f"{tag}_save_tau"])
        ev = np.array([[e[0], e[1], e[2]] for e in tr.events()])
        assert np.array_equal(ev, g[f"{tag}_events"])


@pytest.mark.parametrize("N,dt", [(32, 1e-5), (96, 4e-6)])
def test_oracle_vs_numpy_transcription_and_expm(oracle_mod, N, dt):
    w = syn.make_workload(N)
    Hp = _hprime(oracle_mod, w)
    Hp_np, _ = tn.h_prime(w.S, w.h)
    assert np.abs(Hp - Hp_np).max() / np.abs(Hp).max() < 1e-12
    tau0 = dt / tn.H_BAR
    U = expm(-1j * tau0 * Hp)
    for p in range(2):
        b, k, _, st, tr = oracle_mod.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        log = {}
        b2, k2, _, st2 = tn.propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0, log)
        assert np.abs(b - b2).max() < 1e-13 and np.abs(k - k2).max() < 1e-13
        assert st == st2 and tr.n_matvec_pairs == log["matvec_pairs"]
        assert [(e[0], e[1], e[2]) for e in tr.events()] == [(e[0], e[1], e[2]) for e in log["events"]][:256]
        # the algorithm's own accuracy (1e-8 tolerances): close to the exact propagator, not equal to it
        assert np.abs(U @ w.Psi_ket[:, p] - k).max() < 1e-7
        assert np.abs(U.T @ w.Psi_bra[:, p] - b).max() < 1e-7
        # charge conservation, Taylor.f:104
        assert abs(abs(np.vdot(b, k)) - abs(np.vdot(w.Psi_bra[:, p], w.Psi_ket[:, p]))) < 1e-7


def test_convergence_failure_leaves_psi_untouched(oracle_mod):
    w = syn.make_workload(64)
    Hp = _hprime(oracle_mod, w)
    ok, b, k, C, k_ref, k_exit = oracle_mod.convergence(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], 0.76, 1.0)
    assert not ok and k_exit == 0 and k_ref == 25 or k_ref <= 25
    assert np.array_equal(b, w.Psi_bra[:, 0]) and np.array_equal(k, w.Psi_ket[:, 0])


def test_dzgemv_matches_numpy(oracle_mod):
    rng = np.random.default_rng(5)
    for n in (1, 7, 130, 513):
        A = np.asfortranarray(rng.normal(size=(n, n)))
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        alpha = 0.3 - 1.1j
        assert np.allclose(oracle_mod.dzgemv("N", A, x, alpha), alpha * (A @ x), rtol=1e-13, atol=1e-13)
        assert np.allclose(oracle_mod.dzgemv("T", A, x, alpha), alpha * (A.T @ x), rtol=1e-13, atol=1e-13)


def test_build_huckel_and_x_ij(oracle_mod):
    w = syn.make_workload(64)
    h = oracle_mod.build_huckel(w.IP, w.k_WH, w.V_shift, w.S)
    assert np.allclose(h, tn.x_ij_matrix(w.IP, w.k_WH, w.V_shift) * w.S, rtol=1e-15, atol=0)
    assert np.allclose(h, w.h, rtol=1e-15, atol=0)
    assert np.array_equal(h, h.T)
    assert np.allclose(np.diag(h), w.IP)           # X_ii = IP_i + V_shift_i, S_ii = 1


def test_populations_and_energies(oracle_mod):
    w = syn.make_workload(64)
    frag = w.fragment
    DUAL_bra = np.conj(w.Psi_ket); DUAL_ket = w.Psi_bra
    pops = oracle_mod.populations(frag, DUAL_bra, DUAL_ket, 0.5, 4)
    for p in range(2):
        assert pops[0, p] == 0.5
        assert np.allclose(pops[1:, p], tn.pop_slater(frag, DUAL_bra[:, p], DUAL_ket[:, p], 4), atol=1e-15)
        assert abs(pops[5, p] - 1.0) < 1e-12
    e = oracle_mod.quasiparticle_energies(np.conj(w.Psi_ket), w.Psi_ket, w.h)
    for p in range(2):
        assert np.allclose(e[p], np.conj(w.Psi_ket[:, p]) @ w.h @ w.Psi_ket[:, p], rtol=1e-12)


def test_elhl_step_matches_composition(oracle_mod):
    """ElHl_Chebyshev.f:148-291 == syInvert + syMultiply + Propagation per particle + post-processing."""
    N, dt = 64, 2e-6
    w = syn.make_workload(N)
    st = oracle_mod.ElHlState(w.Psi_bra, w.Psi_ket)
    out = oracle_mod.elhl_step(st, w.S, w.h, dt)
    Hp = _hprime(oracle_mod, w)
    assert np.abs(out["H_prime"] - Hp).max() == 0.0
    Sinv = oracle_mod.sy_invert(w.S)
    for p in range(2):
        b, k, _, save_tau, _ = oracle_mod.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, dt / tn.H_BAR)
        assert np.array_equal(st.Psi_bra[:, p], b) and np.array_equal(st.Psi_ket[:, p], k)
        assert st.save_tau[p] == save_tau
        assert np.array_equal(out["DUAL_bra"][:, p], np.conj(k)) and np.array_equal(out["DUAL_ket"][:, p], b)
        assert np.allclose(out["AO_bra"][:, p], np.conj(Sinv @ b), rtol=1e-12, atol=1e-14)
    assert out["t"] == dt


def test_golden_trajectory(oracle_mod, golden_dir):
    g = np.load(os.path.join(golden_dir, "traj_N64_dt2e-6_100steps.npz"))
    N, dt, n_steps = int(g["N"]), float(g["dt"]), 5        # first 5 of the 20 golden steps keep the CPU suite fast
    pos, species = syn.lattice(N // 4, 1234 + N)
    S0, _ = syn.workload_at(pos, species)
    _, Psi_bra, Psi_ket = syn.packets(S0, N)
    st = oracle_mod.ElHlState(Psi_bra, Psi_ket)
    frag = syn.fragments(N)
    for step in range(n_steps):
        S, h = syn.workload_at(syn.perturb_positions(pos, step), species)
        o = oracle_mod.elhl_step(st, S, h, dt)
        pops = oracle_mod.populations(frag, o["DUAL_bra"], o["DUAL_ket"], o["t"], 4)
        assert np.abs(pops - g["pops"][step]).max() < 1e-13
        assert [tr.n_matvec_pairs for tr in o["traces"]] == list(g["pairs"][step])


# ---------------------------------------------------------------------------------- the real reference, where it compiles
def test_naked_bessel_against_reference_binary(oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    for n in range(2, 25):
        for x in (0.05, 0.3, 0.7596, 2.5):
            assert oracle_mod.naked_bessel(n, x) == oracle_mod.ref_nakedbessel(n, x)


def test_syinvert_symm_against_reference_binary(oracle_mod):
    """xpu_syinvert_/xpu_dsymm_ compiled from /root/reference/GPU_Interface.cpp (CPU mode)."""
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    for N in (32, 128, 256):
        w = syn.make_workload(N)
        ours = oracle_mod.sy_invert(w.S)
        ref_upper = np.triu(oracle_mod.ref_sy_invert_upper(w.S))
        scale = np.abs(ours).max()
        assert np.abs(np.triu(ours) - ref_upper).max() / scale < 1e-12
        assert np.array_equal(ours, ours.T)                       # Matrix_Symmetrize('U'), Matrix_math.f:125-174
        Hp = oracle_mod.sy_multiply(ours, w.h)
        Hp_ref = oracle_mod.ref_dsymm_LU(ours, w.h)
        assert np.abs(Hp - Hp_ref).max() / np.abs(Hp).max() < 1e-13


def test_chebyshev_variant_small_tau(oracle_mod):
    """The un-linked Chebyshev series (Chebyshev_gpu.cpp:524-632) is only valid for tau*rho(H) <~ 1 (no
    rescaling); in that regime it must agree with expm."""
    w = syn.make_workload(32)
    Hp = _hprime(oracle_mod, w)
    rho = np.abs(np.linalg.eigvals(Hp)).max()
    tau = 0.5 / rho
    ok, b, k, C, k_ref, k_exit = oracle_mod.cheb_convergence(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], tau, 1.0)
    assert ok and 2 <= k_exit < k_ref
    U = expm(-1j * tau * Hp)
    assert np.abs(U @ w.Psi_ket[:, 0] - k).max() < 1e-7
    assert np.abs(U.T @ w.Psi_bra[:, 0] - b).max() < 1e-7
    c = oracle_mod.cheb_coefficient(tau)
    from scipy.special import jv
    assert np.allclose(c[0], jv(0, tau)) and np.allclose(c[3], 2 * jv(3, tau) * (-1j) ** 3)


# ---------------------------------------------------------------------------------- Chebyshev mode (product spec)
@pytest.mark.parametrize("name", ["cheb_N64_dt5e-4", "cheb_N128_dt5e-5"])
def test_scaled_chebyshev_oracle_reproduces_golden(oracle_mod, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    Hp = g["H_prime"]; emin, emax = float(g["emin"]), float(g["emax"])
    ebar, de = 0.5 * (emax + emin), 0.5 * (emax - emin)
    for p, tag in enumerate(("el", "hl")):
        b, k, _, st, tr = oracle_mod.cheb_scaled_propagation(Hp, g["Psi_bra0"][:, p], g["Psi_ket0"][:, p],
                                                             float(g["t_init"]), float(g["t_max"]), float(g["tau0"]), ebar, de)
        assert np.abs(b - g[f"{tag}_bra"]).max() < 1e-14 and np.abs(k - g[f"{tag}_ket"]).max() < 1e-14
        assert tr.n_matvec_pairs == int(g[f"{tag}_matvec_pairs"]) and st == float(g[f"{tag}_save_tau"])


def test_scaled_chebyshev_vs_numpy_and_expm(oracle_mod):
    N, dt = 48, 2e-4
    w = syn.make_workload(N)
    Hp = _hprime(oracle_mod, w)
    e = np.linalg.eigvals(Hp).real
    lo, hi = e.min() - 0.05 * (e.max() - e.min()), e.max() + 0.05 * (e.max() - e.min())
    ebar, de = 0.5 * (hi + lo), 0.5 * (hi - lo)
    tau0 = dt / tn.H_BAR
    U = expm(-1j * tau0 * Hp)
    b, k, _, st, tr = oracle_mod.cheb_scaled_propagation(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], 0.0, dt, tau0, ebar, de)
    log = {}
    b2, k2, _, st2 = tn.cheb_propagation(Hp, w.Psi_bra[:, 0].copy(), w.Psi_ket[:, 0].copy(), 0.0, dt, tau0, ebar, de, log)
    assert np.abs(b - b2).max() < 1e-11 and np.abs(k - k2).max() < 1e-11 and st == st2
    assert [(x[0], x[1], x[2]) for x in tr.events()] == [(x[0], x[1], x[2]) for x in log["events"]][:256]
    assert np.abs(U @ w.Psi_ket[:, 0] - k).max() < 1e-7 and np.abs(U.T @ w.Psi_bra[:, 0] - b).max() < 1e-7
    # far fewer dual products than the Taylor series for the same step
    _, _, _, _, tr_t = oracle_mod.propagation(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], 0.0, dt, tau0)
    assert tr.n_matvec_pairs * 5 < tr_t.n_matvec_pairs
    # with ebar = 0, de = 1 the scaled series IS the reference's un-linked series (Chebyshev_gpu.cpp:524-632)
    tau = 0.3 / np.abs(e).max()
    ok, b1, k1, C1, kr1, kx1 = oracle_mod.cheb_convergence(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], tau, 1.0)
    b3, k3, _, _, tr3 = oracle_mod.cheb_scaled_propagation(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], 0.0, tau * tn.H_BAR, tau, 0.0, 1.0)
    assert ok and tr3.events()[0][1] == kx1 and np.array_equal(b1, b3) and np.array_equal(k1, k3)


def test_gpu_variant_transcription_vs_cpu_variant(oracle_mod):
    """SURVEY.md Appendix B in numbers: the numpy transcription of the reference's GPU variant (Taylor_gpu.cpp:334-622,
    oracle/taylor_numpy.py gpu_variant_*; pinned against the reference's own binary on the GPU box by
    tests/test_gpu_reference_gpu.py) and the CPU variant the product follows (Taylor.f) propagate the same packet to
    the same state within the series tolerance, with different tau schedules (the GPU variant sums one term fewer and
    settles on a smaller tau)."""
    from oracle import taylor_numpy as tn
    from dynemol_b200 import synthetic as syn
    N, dt = 128, 5e-6
    w = syn.make_workload(N)
    Hp = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    tau0 = dt / tn.H_BAR
    for p in range(2):
        gb, gk, _, g_save, _ = tn.gpu_variant_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0)
        cb, ck, _, c_save, _ = oracle_mod.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        assert np.abs(gb - cb).max() / np.abs(cb).max() < 2e-7 and np.abs(gk - ck).max() / np.abs(ck).max() < 2e-7
        assert abs(abs(np.vdot(gb, gk)) - 1.0) < 1e-7
        assert 0.0 < g_save <= c_save <= tau0


def test_chebyshev_gpu_variant_transcription_vs_expm(oracle_mod):
    """The numpy transcription of the reference's un-linked Chebyshev GPU driver (Chebyshev_gpu.cpp:347-485,524-643,
    oracle/taylor_numpy.py gpu_variant_cheb_*; pinned against the reference's own binary on the GPU box by
    tests/test_gpu_refgpu_modes.py) agrees with expm and with the CPU-style restatement (ebar = 0, de = 1) within the
    series tolerance; the two differ only in the term test (Idamax element + strict `<` against complex modulus + `>`)."""
    from oracle import taylor_numpy as tn
    N, dt = 96, 2e-6
    w = syn.make_workload(N)
    Hp = _hprime(oracle_mod, w)
    tau0 = dt / tn.H_BAR
    U = expm(-1j * tau0 * Hp)
    for p in range(2):
        gb, gk, _, g_save, _ = tn.gpu_variant_cheb_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0)
        cb, ck, _, c_save, _ = oracle_mod.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0, 0.0, 1.0)
        assert np.abs(gb - cb).max() / np.abs(cb).max() < 2e-7 and np.abs(gk - ck).max() / np.abs(ck).max() < 2e-7
        assert np.abs(U @ w.Psi_ket[:, p] - gk).max() < 2e-7 and np.abs(U.T @ w.Psi_bra[:, p] - gb).max() < 2e-7
        assert 0.0 < g_save <= tau0 and 0.0 < c_save <= tau0


def test_single_expansion_chebyshev_spec_vs_expm(oracle_mod):
    """The product's DYB_MODE_CHEBYSHEV_FULL as stated in numpy (oracle/taylor_numpy.py cheb_full_*): one expansion of
    K ~ R + O(R^(1/3)) terms for a whole 0.5 fs step agrees with expm to rounding (1e-10 is the bar in the GPU test) with
    several times fewer dual products than the reference's chain of order-25 series needs for the same step."""
    N, dt = 96, 5e-4
    w = syn.make_workload(N)
    Hp = _hprime(oracle_mod, w)
    e = np.linalg.eigvals(Hp).real
    lo, hi = e.min() - 0.05 * (e.max() - e.min()), e.max() + 0.05 * (e.max() - e.min())
    ebar, de = 0.5 * (hi + lo), 0.5 * (hi - lo)
    U = expm(-1j * (dt / tn.H_BAR) * Hp)
    for p in range(2):
        b, k, n_terms, ok = tn.cheb_full_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, ebar, de)
        assert ok
        assert np.abs(U @ w.Psi_ket[:, p] - k).max() < 1e-11 and np.abs(U.T @ w.Psi_bra[:, p] - b).max() < 1e-11
        _, _, _, _, tr = oracle_mod.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, dt / tn.H_BAR, ebar, de)
        R = de * dt / tn.H_BAR
        assert R < n_terms < R + 20.0 * R ** (1.0 / 3.0) + 30
        assert 2 * n_terms < tr.n_matvec_pairs, (n_terms, tr.n_matvec_pairs)
