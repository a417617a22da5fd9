#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run from the repo root).

The reference holds no golden vectors for this path, and its
CPU path cannot be built here, so these fixtures are produced by OUR oracle
(oracle/elhl_oracle.cpp) and are only written after three independent checks
pass on the very same inputs:

  1. the numpy transcription (oracle/taylor_numpy.py) reproduces the vectors to
     1e-12 and the decision trace exactly;
  2. scipy.linalg.expm(-i t H'/h_bar) agrees to the 1e-8..1e-7 level the
     reference algorithm itself delivers (its own 1e-8 term/norm tolerances);
  3. where the real reference compiles (oracle/_ref), S^-1 and H' = S^-1 h agree
     with its xpu_syinvert_/xpu_dsymm_ to ~cond(S)*eps.

Fixtures are small (N = 64, 128) so that they can live in git.
"""
from __future__ import annotations

import os
import sys

import numpy as np
from scipy.linalg import expm

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

import oracle                                   # noqa: E402
from oracle import taylor_numpy as tn           # noqa: E402
from dynemol_b200 import synthetic as syn       # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def events_array(tr):
    ev = tr.events()
    return np.array([[e[0], e[1], e[2]] for e in ev], dtype=np.int64), np.array([e[3] for e in ev])


def propagation_case(name, N, dt, t_init=0.0):
    w = syn.make_workload(N)
    Sinv = oracle.sy_invert(w.S)
    Hp = oracle.sy_multiply(Sinv, w.h)
    if oracle.ref_available():
        ref_inv = oracle.ref_sy_invert_upper(w.S)
        assert np.abs(np.triu(ref_inv) - np.triu(Sinv)).max() < 1e-10
        assert np.abs(oracle.ref_dsymm_LU(Sinv, w.h) - Hp).max() < 1e-9
    tau0 = dt / tn.H_BAR
    out = dict(N=N, dt=dt, t_init=t_init, t_max=t_init + dt, tau0=tau0, H_prime=Hp, S=w.S, h=w.h,
               Psi_bra0=w.Psi_bra, Psi_ket0=w.Psi_ket, fragment=w.fragment)
    U = expm(-1j * (dt / tn.H_BAR) * Hp)
    for p, tag in enumerate(("el", "hl")):
        b, k, tau_out, save_tau, tr = oracle.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], t_init, t_init + dt, tau0)
        log = {}
        b2, k2, tau2, st2 = tn.propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), t_init, t_init + dt, tau0, log)
        assert np.abs(b - b2).max() < 1e-12 and np.abs(k - k2).max() < 1e-12, "numpy transcription disagrees"
        n_ev = min(tr.n_events, 256)             # the C trace keeps the first 256 events, counts all
        assert tr.n_events == len(log["events"]), "decision trace length differs"
        assert [(e[0], e[1], e[2]) for e in tr.events()] == [(e[0], e[1], e[2]) for e in log["events"][:n_ev]], "decision trace differs"
        assert log["matvec_pairs"] == tr.n_matvec_pairs
        # exact propagator: ket' = U ket ; bra' = U^T bra  (bra_x_op uses H'^T, Taylor.f:94)
        ek = np.abs(U @ w.Psi_ket[:, p] - k).max(); eb = np.abs(U.T @ w.Psi_bra[:, p] - b).max()
        assert ek < 5e-7 and eb < 5e-7, (ek, eb)
        evk, evt = events_array(tr)
        out.update({f"{tag}_bra": b, f"{tag}_ket": k, f"{tag}_tau_out": tau_out, f"{tag}_save_tau": save_tau,
                    f"{tag}_events": evk, f"{tag}_event_tau": evt, f"{tag}_matvec_pairs": tr.n_matvec_pairs,
                    f"{tag}_substeps": tr.n_substeps, f"{tag}_expm_err": max(ek, eb)})
        print(f"  {name}/{tag}: pairs={tr.n_matvec_pairs} substeps={tr.n_substeps} conv_calls={tr.n_convergence_calls} "
              f"rescale={tr.n_rescale} expm_err={max(ek, eb):.2e}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def chebyshev_case(name, N, dt):
    """Chebyshev mode of the product: the rescaled series (elhl_oracle.cpp orc_cheb_scaled_*), cross-checked with
    the numpy transcription and expm; bounds = exact spectrum widened by 5% (stored in the fixture)."""
    w = syn.make_workload(N)
    Hp = oracle.sy_multiply(oracle.sy_invert(w.S), w.h)
    e = np.linalg.eigvals(Hp).real
    emin, emax = e.min(), e.max(); wd = emax - emin
    emin -= 0.05 * wd; emax += 0.05 * wd
    ebar, de = 0.5 * (emax + emin), 0.5 * (emax - emin)
    tau0 = dt / tn.H_BAR
    out = dict(N=N, dt=dt, t_init=0.0, t_max=dt, tau0=tau0, H_prime=Hp, Psi_bra0=w.Psi_bra, Psi_ket0=w.Psi_ket, emin=emin, emax=emax)
    U = expm(-1j * tau0 * Hp)
    for p, tag in enumerate(("el", "hl")):
        b, k, tau_out, save_tau, tr = oracle.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0, ebar, de)
        log = {}
        b2, k2, _, st2 = tn.cheb_propagation(Hp, w.Psi_bra[:, p].copy(), w.Psi_ket[:, p].copy(), 0.0, dt, tau0, ebar, de, log)
        assert np.abs(b - b2).max() < 1e-11 and np.abs(k - k2).max() < 1e-11, (np.abs(b - b2).max(), np.abs(k - k2).max())
        assert tr.n_events == len(log["events"]) and log["matvec_pairs"] == tr.n_matvec_pairs
        assert [(e_[0], e_[1], e_[2]) for e_ in tr.events()] == [(e_[0], e_[1], e_[2]) for e_ in log["events"][:256]]
        ek = np.abs(U @ w.Psi_ket[:, p] - k).max(); eb = np.abs(U.T @ w.Psi_bra[:, p] - b).max()
        assert ek < 5e-7 and eb < 5e-7, (ek, eb)
        evk, evt = events_array(tr)
        out.update({f"{tag}_bra": b, f"{tag}_ket": k, f"{tag}_save_tau": save_tau, f"{tag}_events": evk, f"{tag}_event_tau": evt,
                    f"{tag}_matvec_pairs": tr.n_matvec_pairs, f"{tag}_substeps": tr.n_substeps, f"{tag}_expm_err": max(ek, eb)})
        print(f"  {name}/{tag}: pairs={tr.n_matvec_pairs} substeps={tr.n_substeps} conv_calls={tr.n_convergence_calls} "
              f"rescale={tr.n_rescale} expm_err={max(ek, eb):.2e}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def trajectory_case(name, N, dt, n_steps):
    """ElHl_Chebyshev.f:148-291 over n_steps nuclear steps with moving nuclei;
    per-step fragment populations (data_output.f:87-147) are the observable."""
    pos, species = syn.lattice(N // 4, 1234 + N)
    S0, h0 = syn.workload_at(pos, species)
    C, Psi_bra, Psi_ket = syn.packets(S0, N)
    frag = syn.fragments(N)
    st = oracle.ElHlState(Psi_bra, Psi_ket)
    pops = np.zeros((n_steps, 6, 2)); pairs = np.zeros((n_steps, 2), dtype=np.int64)
    save_taus = np.zeros((n_steps, 2))
    for step in range(n_steps):
        S, h = syn.workload_at(syn.perturb_positions(pos, step), species)
        o = oracle.elhl_step(st, S, h, dt)
        pops[step] = oracle.populations(frag, o["DUAL_bra"], o["DUAL_ket"], o["t"], 4)
        pairs[step] = [tr.n_matvec_pairs for tr in o["traces"]]
        save_taus[step] = st.save_tau
    assert np.all(np.abs(pops[:, 5, :] - 1.0) < 1e-6), "total population not conserved"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), N=N, dt=dt, n_steps=n_steps, pops=pops, pairs=pairs,
                        save_tau=save_taus, Psi_bra_final=st.Psi_bra, Psi_ket_final=st.Psi_ket,
                        AO_bra_final=o["AO_bra"], H_prime_final=o["H_prime"])
    print(f"  {name}: {n_steps} steps, pairs/step el={pairs[:, 0].mean():.1f} hl={pairs[:, 1].mean():.1f}, "
          f"pop total dev={np.abs(pops[:, 5, :] - 1).max():.2e}")


if __name__ == "__main__":
    oracle.build()
    print("golden fixtures ->", OUT)
    propagation_case("prop_N64_dt5e-6", 64, 5e-6)        # first-loop shrink + steady loop + last-substep shrink
    propagation_case("prop_N64_dt5e-4", 64, 5e-4)        # + the `rescaling tau` branch (Taylor.f:108-113)
    propagation_case("prop_N128_dt2e-5", 128, 2e-5)
    trajectory_case("traj_N64_dt2e-6_100steps", 64, 2e-6, 100)   # north_star: populations within 1e-9 over 100 nuclear steps
    chebyshev_case("cheb_N64_dt5e-4", 64, 5e-4)          # dt = 0.5 fs (BASELINE config), ~750 terms instead of ~18000
    chebyshev_case("cheb_N128_dt5e-5", 128, 5e-5)
