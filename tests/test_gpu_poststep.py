"""GPU suite: the rows next to the hot path: QuasiParticleEnergies and the diabatic-Ehrenfest kernel
on the device with the resident H', and the LU fallback of the H' formation for a non-positive-definite S."""
import numpy as np
import pytest

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import api as a
    assert a.device_count() > 0
    return a


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_quasiparticle_energies_match_oracle(api, oracle_mod):
    """ElHl_Chebyshev.f:329-371 with AO_bra = conj(S^-1 Psi_bra), AO_ket = Psi_ket (:274-276), after a propagation so that
    the packets are genuinely complex."""
    N, dt = 256, 2e-6
    w = syn.make_workload(N)
    P = api.Propagator(N)
    P.form_hprime(w.S, w.h, want_hprime=False)
    P.set_packets(w.Psi_bra, w.Psi_ket)
    P.propagate(0.0, dt, dt / api.H_BAR)
    bra, ket = P.get_packets()
    Sinv = oracle_mod.sy_invert(w.S)
    AO_bra = np.conj(Sinv @ bra)
    ref = oracle_mod.quasiparticle_energies(AO_bra, ket, w.h)
    got = P.quasiparticle_energies()
    assert relerr(got, ref) < 1e-10
    assert relerr(np.conj(P.ao_bra()), AO_bra) < 1e-10
    P.close()


def test_ehrenfest_kernel_matches_reference_formula(api):
    """K = X o A - H' A (diabatic-Ehren.f:115-119), native entry with the resident H' and the legacy symbol."""
    N = 384
    w = syn.make_workload(N)
    rng = np.random.default_rng(8)
    rho = rng.normal(size=(N, N))
    A = np.asfortranarray(0.5 * (rho + rho.T))                      # A_ad_nd = (rho + rho^T)/2, diabatic-Ehren.f:108
    X = np.asfortranarray(syn.x_matrix(w.IP, w.k_WH, w.V_shift))    # X_ij, hamiltonians.f:33-63
    P = api.Propagator(N)
    Hp = P.form_hprime(w.S, w.h)
    ref = X * A - Hp @ A
    assert relerr(P.ehrenfest_kernel(A, X), ref) < 1e-12
    assert relerr(api.legacy_ehrenfestkernel(Hp, A, X), ref) < 1e-12
    P.close()
    api.gpu_finalize()


def test_ehrenfest_kernel2_builds_the_density_on_the_device(api):
    """ehrenfestkernel2_gpu_ (Taylor_gpu.cpp:801-873): rho(i,j) = Re{ket(j,1) bra(i,1)} - Re{ket(j,2) bra(i,2)}
    (calculate_rho, diabatic-Ehren.f:107), A = (rho + rho^T)/2, K = X o A - H' A -- from the AO packets alone."""
    N = 320
    w = syn.make_workload(N)
    rng = np.random.default_rng(9)
    bra = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    ket = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    X = np.asfortranarray(syn.x_matrix(w.IP, w.k_WH, w.V_shift))
    Hp = np.asfortranarray(np.linalg.solve(w.S, w.h))
    rho = np.real(np.outer(bra[:, 0], ket[:, 0])) - np.real(np.outer(bra[:, 1], ket[:, 1]))      # rho[i, j]
    A = 0.5 * (rho + rho.T)
    ref = X * A - Hp @ A
    K2 = api.legacy_ehrenfestkernel2(bra, ket, Hp, X)
    assert relerr(K2, ref) < 1e-12
    assert relerr(K2, api.legacy_ehrenfestkernel(Hp, np.asfortranarray(A), X)) < 1e-13
    api.gpu_finalize()


def test_formation_lu_fallback_for_indefinite_overlap(api):
    """The reference factorises S with Bunch-Kaufman (CPU) or LU (GPU): an S that is symmetric but not positive
    definite must still give H' = S^-1 h (here: Cholesky fails -> LU with partial pivoting)."""
    N = 128
    w = syn.make_workload(N)
    S = w.S.copy()
    S[5, 5] = -0.7; S[40, 40] = -1.3                                  # symmetric, indefinite, well conditioned enough
    assert np.linalg.eigvalsh(S).min() < 0
    P = api.Propagator(N)
    Hp = P.form_hprime(np.asfortranarray(S), w.h)
    assert relerr(Hp, np.linalg.solve(S, w.h)) < 1e-10
    P.set_packets(w.Psi_bra, w.Psi_ket)
    assert relerr(P.ao_bra(), np.linalg.solve(S, w.Psi_bra)) < 1e-10
    P.close()


def test_build_huckel_on_device(api, oracle_mod):
    """Build_Huckel (ElHl_Chebyshev.f:296-323) on the device: same H' as uploading the host-built h."""
    N = 256
    w = syn.make_workload(N)
    P = api.Propagator(N)
    Hp_a = P.form_hprime(w.S, w.h)
    Hp_b = P.form_hprime_from_overlap(w.S, w.IP, w.k_WH, w.V_shift)
    assert relerr(Hp_b, Hp_a) < 1e-13
    # V_shift and k_WH really enter: perturb them and compare with the oracle's Build_Huckel
    rng = np.random.default_rng(2)
    V = rng.normal(scale=0.3, size=N); K = w.k_WH + rng.uniform(-0.2, 0.2, size=N)
    h2 = oracle_mod.build_huckel(w.IP, K, V, w.S)
    Hp_c = P.form_hprime_from_overlap(w.S, w.IP, K, V)
    assert relerr(Hp_c, np.linalg.solve(w.S, h2)) < 1e-11
    P.close()
