"""GPU suite: trajectory-level and full-size parity against the CPU oracle on the paths the small golden fixtures do not
reach ("no trajectory-level parity on the streaming (TMA) path and none at config-2 size", "N=16384 has no
oracle comparison for propagate").

  * 100 nuclear steps with moving nuclei at N = 2304 (twice: the launch-per-term path -- TMA dual product + fused epilogue --
    and the streamed one-launch series kernel of csrc/mid.cuh that DYB_SERIES_AUTO selects there, with the chained steady
    loop) and at N = 900 (BASELINE config 2 size, shared-memory-resident series kernel with the chained steady loop): per step the
    product forms H' = S^-1 h on the device from the host-built S, h, propagates electron and hole with the carried-over
    tau of ElHl_Chebyshev.f:174-187 and reduces the fragment populations on the device; the oracle propagates the same
    packets with the same H' (Taylor.f:35-219) and reduces them on the host (data_output.f:242-263).
    Bar (north_star): populations within 1e-9 over the 100 steps, wavepackets within 1e-10 at the end, identical
    decision traces at every step.
  * one short Taylor step at N = 16384 (BASELINE config 3) against the oracle at full size.
"""
import numpy as np
import pytest

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

REL_TOL = 1e-10
POP_TOL = 1e-9
H_BAR = 6.58264e-4


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import api as a
    assert a.device_count() > 0, "GPU tests need a CUDA device"
    return a


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def events3(tr):
    return [(e[0], e[1], e[2]) for e in tr.events()]


@pytest.mark.parametrize("N,dt,force,series", [(2304, 5e-7, 1, 1), (2304, 5e-7, 0, 5), (900, 2e-6, 0, 3)])
def test_hundred_step_trajectory_against_oracle(api, oracle_mod, N, dt, force, series):
    n_steps = 100
    pos, species = syn.lattice(N // 4, 1234 + N)
    S0, _ = syn.workload_at(pos, species)
    _, Psi_bra, Psi_ket = syn.packets(S0, N)
    frag = syn.fragments(N)
    P = api.Propagator(N)
    if force:
        P.set_series_kernel(force)
    # 1: launch per term (TMA path, forced), 5: streamed one-launch kernel (AUTO), 3: resident series kernel (AUTO)
    assert P.info()["series_kernel"] == series
    P.set_packets(Psi_bra, Psi_ket)
    o_bra = Psi_bra.copy(order="F"); o_ket = Psi_ket.copy(order="F")
    tau_max = dt / H_BAR
    save_tau = np.zeros(2); o_save = np.zeros(2); t = 0.0
    worst_pop = 0.0; total_pairs = 0; n_sub = 0
    for step in range(n_steps):
        it = step + 2                                            # Chebyshev_driver.f:106 increments before the call
        S, h = syn.workload_at(syn.perturb_positions(pos, step), species)
        Hp = P.form_hprime(S, h)                                 # a2+a3 on the device; the oracle gets the same H'
        tau = np.full(2, tau_max) if step == 0 else np.minimum(tau_max, 1.15 * save_tau)   # ElHl_Chebyshev.f:182-184
        o_tau = np.full(2, tau_max) if step == 0 else np.minimum(tau_max, 1.15 * o_save)
        t_max = dt * 1 * (it - 1)                                # ElHl_Chebyshev.f:176
        save_tau, traces = P.propagate(t, t_max, tau)
        for p in range(2):
            b, k, _, st, tr = oracle_mod.propagation(Hp, o_bra[:, p], o_ket[:, p], t, t_max, o_tau[p])
            o_bra[:, p] = b; o_ket[:, p] = k; o_save[p] = st
            assert events3(traces[p]) == events3(tr), f"step {step} particle {p}"
            assert save_tau[p] == st
            total_pairs += tr.n_matvec_pairs; n_sub += tr.n_substeps
        t = t + dt
        pops = P.populations(frag, 4, t)                         # DUAL_bra = conj(ket), DUAL_ket = bra, on the device
        o_pops = oracle_mod.populations(frag, np.conj(o_ket), o_bra, t, 4)
        worst_pop = max(worst_pop, np.abs(pops - o_pops).max())
        assert np.abs(pops - o_pops).max() < POP_TOL, f"step {step}"
    bra, ket = P.get_packets()
    assert relerr(bra, o_bra) < REL_TOL and relerr(ket, o_ket) < REL_TOL
    for p in range(2):
        assert abs(abs(np.vdot(bra[:, p], ket[:, p])) - 1.0) < 1e-6
    print(f"trajectory N={N}: {n_steps} steps, {total_pairs} oracle matvec pairs, {n_sub} steady sub-steps, "
          f"worst population error {worst_pop:.2e}, final packets {relerr(bra, o_bra):.2e}/{relerr(ket, o_ket):.2e}")
    P.close()


def test_headline_size_step_against_oracle(api, oracle_mod):
    """N = 16384 (BASELINE config 3): H' = S^-1 h formed on the device from the synthetic EHT S, h; one Taylor step short
    enough for the first Convergence to succeed within a dozen terms, electron and hole, against the CPU oracle run at
    full size on the box's host cores: identical decisions, wavepackets within 1e-10."""
    import torch
    N = 16384
    S_t, h_t, _ = syn.make_S_h_torch(N, torch.device("cuda", 0))
    P = api.Propagator(N)
    P.form_hprime_device(S_t.data_ptr(), N, h_t.data_ptr(), N)
    w = 64
    C = np.zeros((N, 2)); C[0:w, 0] = np.random.default_rng(42).normal(size=w); C[w:2 * w, 1] = np.random.default_rng(43).normal(size=w)
    Ct = torch.tensor(C, device="cuda")
    SC = S_t @ Ct
    nrm = torch.sqrt((Ct * SC).sum(0))
    Psi_ket = np.asfortranarray((Ct / nrm).cpu().numpy().astype(np.complex128))
    Psi_bra = np.asfortranarray((SC / nrm).cpu().numpy().astype(np.complex128))
    del S_t, h_t, SC, Ct
    torch.cuda.empty_cache()
    Hp = P.download_hprime()
    P.set_packets(Psi_bra, Psi_ket)
    dt = 1e-7; tau0 = dt / H_BAR
    save_tau, traces = P.propagate(0.0, dt, tau0)
    gb, gk = P.get_packets()
    oracle_mod.use_all_host_threads()
    for p in range(2):
        b, k, _, st, tr = oracle_mod.propagation(Hp, Psi_bra[:, p], Psi_ket[:, p], 0.0, dt, tau0)
        assert tr.n_matvec_pairs <= 64, "the step was meant to be short (CPU oracle at full size)"
        assert events3(traces[p]) == events3(tr) and save_tau[p] == st
        assert relerr(gb[:, p], b) < REL_TOL and relerr(gk[:, p], k) < REL_TOL
    P.close()
