"""GPU suite (needs >= 2 GPUs; `gpurun --gpus 2 -- pytest tests/test_gpu_team.py -m gpu`): single-process multi-GPU.

A dyb_team drives P row-sharded contexts from one host thread per device inside every call (csrc/team.cu), which is what
lets the ONE-process Fortran caller (ElHl_Chebyshev_GPU.f:269-272) reach several GPUs.  Bars: the legacy
symbol with DYNEMOL_B200_GPUS=P reproduces the CPU oracle's nuclear step (identical tau, wavepackets 1e-10, H' 1e-11),
and the team's results equal the single-GPU ones bit for bit in the decisions and to rounding in the vectors."""
import os

import numpy as np
import pytest

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
H_BAR = 6.58264e-4
REL_TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import api as a
    if a.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    return a


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def events3(tr):
    return [(e[0], e[1], e[2]) for e in tr.events()]


def team_sizes(api):
    return [p for p in (2, 4, 8) if p <= api.device_count()]


def test_team_taylor_step_matches_oracle(api, oracle_mod):
    N, dt = 2048, 5e-7
    w = syn.make_workload(N)
    Hp_or = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    tau0 = dt / H_BAR
    frag = syn.fragments(N)
    for P in team_sizes(api):
        T = api.Team(N, P)
        Hp = T.form_hprime(w.S, w.h)
        assert relerr(Hp, Hp_or) < 1e-11
        T.upload_hprime(Hp_or)                                   # same H' on both sides for the recursion
        T.set_packets(w.Psi_bra, w.Psi_ket)
        save_tau, traces = T.propagate(0.0, dt, tau0)
        bra, ket = T.get_packets()
        pops = T.populations(frag, 4, dt)
        erg = T.quasiparticle_energies()
        for p in range(2):
            b, k, _, st, tr = oracle_mod.propagation(Hp_or, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
            assert events3(traces[p]) == events3(tr) and save_tau[p] == st
            assert relerr(bra[:, p], b) < REL_TOL and relerr(ket[:, p], k) < REL_TOL
            assert abs(erg[p] - np.vdot(b, Hp_or @ k)) < 1e-9 * abs(erg[p])
        o_pops = oracle_mod.populations(frag, np.conj(ket), bra, dt, 4)
        assert np.abs(pops - o_pops).max() < 1e-12
        T.close()


def test_legacy_symbol_on_a_team(api, oracle_mod):
    """propagationelhl2_gpucaller_ with DYNEMOL_B200_GPUS=P: host S, h in; H', AO_bra, packets out -- the whole nuclear
    step of ElHl_Chebyshev.f:148-291 against the oracle, N = 8192 on 2 GPUs (and 4 / 8 when the box has them)."""
    N, dt = 8192, 2e-7
    import torch
    S_t, h_t, _ = syn.make_S_h_torch(N, torch.device("cuda", 0))
    S = np.asfortranarray(S_t.cpu().numpy()); h = np.asfortranarray(h_t.cpu().numpy())
    del S_t, h_t
    torch.cuda.empty_cache()
    _, Psi_bra, Psi_ket = syn.packets(S, N)
    tau0 = dt / H_BAR
    oracle_mod.use_all_host_threads()
    P1 = api.Propagator(N)
    Hp1 = P1.form_hprime(S, h)
    P1.close()
    ref = []
    for p in range(2):
        ref.append(oracle_mod.propagation(Hp1, Psi_bra[:, p], Psi_ket[:, p], 0.0, dt, tau0))
    try:
        # formation: distributed (factor on the first GPU, column-block solves on every GPU, block exchange) by default;
        # "single" = whole solve on the first GPU, row blocks scattered -- both must give the same H'
        for P, form in [(p, "distributed") for p in team_sizes(api)] + [(2, "single")]:
            os.environ["DYNEMOL_B200_GPUS"] = str(P)
            os.environ["DYNEMOL_B200_TEAM_FORM"] = form
            out = api.legacy_propagationelhl(S, h, Psi_bra, Psi_ket, 0.0, dt, tau0)
            assert relerr(out["H_prime"], Hp1) < 1e-12
            Sx = lambda z: (S @ z.real) + 1j * (S @ z.imag)
            for p in range(2):
                b, k, _, st, tr = ref[p]
                assert out["save_tau"][p] == st
                assert relerr(out["PSI_bra"][:, p], b) < REL_TOL and relerr(out["PSI_ket"][:, p], k) < REL_TOL
                assert np.abs(Sx(out["AO_bra"][:, p]) - out["PSI_bra"][:, p]).max() < 1e-10      # AO_bra = S^-1 PSI_bra
            api.gpu_finalize()
    finally:
        os.environ.pop("DYNEMOL_B200_GPUS", None); os.environ.pop("DYNEMOL_B200_TEAM_FORM", None)
        api.gpu_finalize()


def test_team_chebyshev_modes_match_single_gpu(api, oracle_mod):
    """Sharded Lanczos bounds and both Chebyshev modes on a team against the same calls on one GPU."""
    N, dt = 1024, 2e-4
    w = syn.make_workload(N)
    Hp = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    tau0 = dt / H_BAR
    P1 = api.Propagator(N)
    P1.set_series_kernel("term")
    P1.upload_hprime(Hp)
    single = {}
    for mode in (api.MODE_CHEBYSHEV, api.MODE_CHEBYSHEV_FULL):
        P1.set_packets(w.Psi_bra, w.Psi_ket)
        bounds = P1.estimate_spectral_bounds(24, 0.05)
        save, tr = P1.propagate(0.0, dt, tau0, mode=mode)
        single[mode] = (bounds, save, [events3(t) for t in tr], P1.get_packets())
    P1.close()
    for P in team_sizes(api):
        T = api.Team(N, P)
        T.upload_hprime(Hp)
        for mode in (api.MODE_CHEBYSHEV, api.MODE_CHEBYSHEV_FULL):
            T.set_packets(w.Psi_bra, w.Psi_ket)
            bounds = T.estimate_spectral_bounds(24, 0.05)
            b1, s1, e1, (bra1, ket1) = single[mode]
            assert abs(bounds[0] - b1[0]) < 1e-9 * (b1[1] - b1[0]) and abs(bounds[1] - b1[1]) < 1e-9 * (b1[1] - b1[0])
            T.set_spectral_bounds(*b1)                           # identical interval: identical decisions
            save, tr = T.propagate(0.0, dt, tau0, mode=mode)
            bra, ket = T.get_packets()
            assert [events3(t) for t in tr] == e1 and np.array_equal(save, s1)
            assert relerr(bra, bra1) < 1e-12 and relerr(ket, ket1) < 1e-12
        T.close()
