"""GPU suite: the product's reference-GPU parity modes against THE REFERENCE'S OWN BINARIES run on the same B200.

DYB_MODE_TAYLOR_REFGPU reproduces Taylor_gpu.cpp:334-480,511-622 decision for decision (one term fewer per series, raw
powers H'^k psi with c_k applied in the sum, cublasIdamax-style term test with a strict `<`), DYB_MODE_CHEBYSHEV_REFGPU
the un-linked Chebyshev_gpu.cpp:347-485,524-643 (no spectral rescaling).  oracle/_ref/libref_taylor_gpu.so and
libref_chebyshev_gpu.so are those files compiled in place from /root/reference (oracle/Makefile); they export the very
legacy symbols the product replaces, so both sides are called through the same Fortran ABI with the same host buffers.

Bar (north_star): identical converged tau (`save_tau`, i.e. the same sequence of accept/reject decisions of the first
Convergence loop), wavepackets within 1e-10 relative after the full nuclear step.  This is parity against reference
code that actually ran -- the CPU-variant modes are pinned by the oracle restatement instead (tests/test_gpu_parity.py).
"""
import os

import numpy as np
import pytest

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
H_BAR = 6.58264e-4
TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import api as a
    assert a.device_count() > 0
    return a


@pytest.fixture(scope="module")
def refgpu(oracle_mod):
    if not oracle_mod.ref_gpu_available():
        pytest.skip("oracle/_ref/libref_taylor_gpu.so not built (needs /root/reference at build time)")
    return oracle_mod


@pytest.fixture(scope="module")
def refcheb(oracle_mod):
    if not oracle_mod.ref_cheb_gpu_available():
        pytest.skip("oracle/_ref/libref_chebyshev_gpu.so not built (needs /root/reference at build time)")
    return oracle_mod


@pytest.fixture
def mode_env():
    """Selects the propagator of the legacy symbols (they have no mode argument: DYNEMOL_B200_MODE)."""
    def set_mode(m):
        os.environ["DYNEMOL_B200_MODE"] = m
    yield set_mode
    os.environ.pop("DYNEMOL_B200_MODE", None)


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def hprime(api, w):
    P = api.Propagator(w.S.shape[0])
    Hp = P.form_hprime(w.S, w.h)
    P.close()
    return Hp


@pytest.mark.parametrize("N,dt", [(256, 5e-6), (1024, 1e-6), (4096, 4e-7)])
def test_taylor_refgpu_propagation_symbol(api, refgpu, mode_env, N, dt):
    """propagation_gpucaller_ (Taylor_gpu.cpp:295-330), H' given: N=256, 1024 run through the shared-memory-resident
    series kernel, N=4096 through the streaming (TMA) dual product + fused epilogue."""
    w = syn.make_workload(N)
    Hp = hprime(api, w)
    tau0 = dt / H_BAR
    mode_env("taylor_refgpu")
    for p in range(2):
        rb, rk, r_save = refgpu.ref_gpu_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        ob, ok_, o_save = api.legacy_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0)
        assert o_save == pytest.approx(r_save, rel=1e-13), "first Convergence loop settled on another tau"
        assert relerr(ob, rb) < TOL and relerr(ok_, rk) < TOL
    api.gpu_finalize()


def test_taylor_refgpu_decisions_match_transcription(api, refgpu):
    """The same mode through the native API, with its decision trace: number of rescales and the converged tau agree with
    the line-by-line numpy transcription of Taylor_gpu.cpp (which itself reproduces the reference binary,
    tests/test_gpu_reference_gpu.py::test_reference_gpu_algorithm_as_transcribed)."""
    from oracle import taylor_numpy as tn
    N, dt = 256, 2e-5
    w = syn.make_workload(N)
    Hp = hprime(api, w)
    tau0 = dt / H_BAR
    for kind in ("resident", "term"):
        P = api.Propagator(N)
        P.set_series_kernel(kind)
        P.upload_hprime(Hp)
        P.set_packets(w.Psi_bra, w.Psi_ket)
        save, tr = P.propagate(0.0, dt, tau0, mode=api.MODE_TAYLOR_REFGPU)
        b, k = P.get_packets()
        P.close()
        for p in range(2):
            nb, nk, n_tau, n_save, n_resc = tn.gpu_variant_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0)
            assert save[p] == pytest.approx(n_save, rel=1e-14)
            assert tr[p].n_rescale == n_resc
            assert tr[p].final_tau == pytest.approx(n_tau, rel=1e-12)
            assert relerr(b[:, p], nb) < TOL and relerr(k[:, p], nk) < TOL


def test_taylor_refgpu_elhl_symbol(api, refgpu, mode_env):
    """propagationelhl_gpucaller_ (Taylor_gpu.cpp:634-736), the symbol ElHl_Chebyshev_GPU.f:269 calls: host S, h in;
    H', AO_bra and the propagated packets out.  The reference inverts S by LU and multiplies (GPU_Interface.cpp:910-929,
    Taylor_gpu.cpp:687-696); the product solves with the Cholesky factor -- same H' to O(cond(S) eps)."""
    N, dt = 768, 2e-6
    w = syn.make_workload(N)
    tau0 = dt / H_BAR
    mode_env("taylor_refgpu")
    for p in range(2):
        r_Hp, r_ao, rb, rk, r_save = refgpu.ref_gpu_propagationelhl(w.S, w.h, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        o = api.legacy_propagationelhl(w.S, w.h, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0)
        assert relerr(o["H_prime"], r_Hp) < 1e-11
        assert o["save_tau"][0] == pytest.approx(r_save, rel=1e-13)
        assert relerr(o["PSI_bra"], rb) < TOL and relerr(o["PSI_ket"], rk) < TOL
        assert relerr(o["AO_bra"], r_ao) < TOL
    # the batched symbol serves both particles with the same passes: identical to the two per-particle calls
    o2 = api.legacy_propagationelhl(w.S, w.h, w.Psi_bra.copy(), w.Psi_ket.copy(), 0.0, dt, tau0)
    for p in range(2):
        r_Hp, r_ao, rb, rk, r_save = refgpu.ref_gpu_propagationelhl(w.S, w.h, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        assert o2["save_tau"][p] == pytest.approx(r_save, rel=1e-13)
        assert relerr(o2["PSI_bra"][:, p], rb) < TOL and relerr(o2["PSI_ket"][:, p], rk) < TOL and relerr(o2["AO_bra"][:, p], r_ao) < TOL
    api.gpu_finalize()


@pytest.mark.parametrize("N,dt", [(256, 2e-6), (1024, 5e-7), (2304, 4e-7)])
def test_chebyshev_refgpu_propagation_symbol(api, refcheb, mode_env, N, dt):
    """The reference's un-linked Chebyshev/Bessel driver (Chebyshev_gpu.cpp, makefile:184-186) against
    DYB_MODE_CHEBYSHEV_REFGPU = the product's Chebyshev recurrence with Ebar = 0, Delta E = 1 and the GPU-style term test.
    Small tau only: the reference does not rescale H' (SURVEY.md a9)."""
    w = syn.make_workload(N)
    Hp = hprime(api, w)
    tau0 = dt / H_BAR
    mode_env("chebyshev_refgpu")
    for p in range(2):
        rb, rk, r_save = refcheb.ref_gpu_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0, chebyshev=True)
        ob, ok_, o_save = api.legacy_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0)
        assert o_save == pytest.approx(r_save, rel=1e-13)
        assert relerr(ob, rb) < TOL and relerr(ok_, rk) < TOL
        assert abs(abs(np.vdot(ob, ok_)) - 1.0) < 1e-7
    api.gpu_finalize()


def test_reference_chebyshev_gpu_algorithm_as_transcribed(refcheb):
    """The numpy transcription of Chebyshev_gpu.cpp (oracle/taylor_numpy.py: gpu_variant_cheb_*) reproduces the reference
    binary: what DYB_MODE_CHEBYSHEV_REFGPU implements is the reference's algorithm, not a reading of it."""
    from oracle import taylor_numpy as tn
    N, dt = 256, 2e-6
    w = syn.make_workload(N)
    Hp = refcheb.sy_multiply(refcheb.sy_invert(w.S), w.h)
    tau0 = dt / H_BAR
    for p in range(2):
        rb, rk, r_save = refcheb.ref_gpu_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0, chebyshev=True)
        nb, nk, _, n_save, _ = tn.gpu_variant_cheb_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0)
        assert n_save == pytest.approx(r_save, rel=1e-14)
        assert relerr(nb, rb) < 1e-11 and relerr(nk, rk) < 1e-11
