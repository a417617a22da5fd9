"""GPU suite: the weak xPU_* dispatchers of the propagator path (GPU_Interface.cpp:129-158; Matrix_math.f routes syInvert,
syMultiply and bra_x_op / op_x_ket through them) against THE REFERENCE'S OWN GPU_Interface.cpp compiled in CPU mode
(oracle/_ref/libref_xpu_cpu.so on top of OpenBLAS) and against numpy."""
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


def test_xpu_syinvert_and_dsymm_against_reference_binary(api, oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built")
    N = 300
    w = syn.make_workload(N)
    ref_inv = oracle_mod.ref_sy_invert_upper(w.S)                  # only the upper triangle is meaningful (dsytri)
    inv, info = api.xpu_syinvert(w.S, "U")
    assert info == 0
    iu = np.triu_indices(N)
    assert relerr(inv[iu], ref_inv[iu]) < 1e-11
    assert np.array_equal(inv, inv.T), "both triangles are returned"
    assert relerr(inv @ w.S, np.eye(N)) < 1e-10
    ref_Hp = oracle_mod.ref_dsymm_LU(ref_inv, w.h)
    Hp = api.xpu_dsymm(inv, w.h, "L", "U")
    assert relerr(Hp, ref_Hp) < 1e-11
    # beta != 0 and the lower triangle
    C0 = np.asfortranarray(np.random.default_rng(1).normal(size=(N, N)))
    out = api.xpu_dsymm(np.tril(inv), w.h, "L", "L", alpha=0.5, beta=2.0, Cin=C0)
    assert relerr(out, 0.5 * inv @ w.h + 2.0 * C0) < 1e-12


def test_xpu_syinvert_indefinite_matrix(api):
    rng = np.random.default_rng(3)
    N = 64
    A = rng.normal(size=(N, N)); A = A + A.T                      # symmetric, indefinite: Cholesky fails, LU route
    inv, info = api.xpu_syinvert(A, "U")
    assert info == 0 and relerr(inv @ A, np.eye(N)) < 1e-9


def test_xpu_dzgemv_matches_numpy(api, oracle_mod):
    rng = np.random.default_rng(4)
    N = 257
    A = np.asfortranarray(rng.normal(size=(N, N)))
    x = rng.normal(size=N) + 1j * rng.normal(size=N)
    y0 = rng.normal(size=N) + 1j * rng.normal(size=N)
    alpha = 0.3 - 1.1j
    assert relerr(api.xpu_dzgemv("N", A, x, alpha), alpha * (A @ x)) < 1e-13
    assert relerr(api.xpu_dzgemv("T", A, x, alpha), alpha * (A.T @ x)) < 1e-13
    assert relerr(api.xpu_dzgemv("N", A, x, alpha, beta=2.0 + 1.0j, y=y0), alpha * (A @ x) + (2.0 + 1.0j) * y0) < 1e-13
    # the oracle's statement of the MKL routine the CPU path calls (Matrix_math.f:221-301)
    assert relerr(api.xpu_dzgemv("T", A, x, alpha), oracle_mod.dzgemv("T", A, x, alpha)) < 1e-13
