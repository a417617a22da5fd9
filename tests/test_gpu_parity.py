"""GPU suite (run on the B200 box: `pytest -m gpu`).  Every check goes through the C ABI
(dynemol_b200/api.py -> libdynemol_b200.so) and compares with the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): wavepacket coefficients within 1e-10 relative after a full
nuclear step; fragment populations within 1e-9 over a trajectory.  Decision traces (tau schedule,
exit index of every Convergence call, sub-step count) must be IDENTICAL, not merely close."""
import os

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


# ----------------------------------------------------------------------------- kernel level: one dual product
@pytest.mark.parametrize("kernel", ["tma", "ldg"])
@pytest.mark.parametrize("N", [1, 2, 63, 257, 1000, 2049, 4100])
def test_dual_matvec_matches_oracle(api, oracle_mod, N, kernel):
    """H' x_ket ('N') and H'^T x_bra ('T') for el+hole in one pass; odd N exercises the padding rules."""
    rng = np.random.default_rng(100 + N)
    H = np.asfortranarray(rng.normal(size=(N, N)))
    xb = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    xk = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    P = api.Propagator(N, kernel=api.KERNEL_TMA if kernel == "tma" else api.KERNEL_LDG)
    P.upload_hprime(H)
    yb, yk = P.dual_matvec(xb, xk)
    for p in range(2):
        assert relerr(yb[:, p], oracle_mod.dzgemv("T", H, xb[:, p])) < 1e-13
        assert relerr(yk[:, p], oracle_mod.dzgemv("N", H, xk[:, p])) < 1e-13
    # single-particle call: hole slots must come back untouched by the electron
    yb1, yk1 = P.dual_matvec(xb[:, 0], xk[:, 0])
    assert np.array_equal(yb1[:, 0], yb[:, 0]) and np.array_equal(yk1[:, 0], yk[:, 0])
    P.close()


def test_kernels_are_deterministic_and_agree(api):
    N = 3000
    rng = np.random.default_rng(7)
    H = np.asfortranarray(rng.normal(size=(N, N)))
    xb = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    xk = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    P = api.Propagator(N)
    P.upload_hprime(H)
    P.set_kernel(api.KERNEL_TMA)
    a1 = P.dual_matvec(xb, xk); a2 = P.dual_matvec(xb, xk)
    assert np.array_equal(a1[0], a2[0]) and np.array_equal(a1[1], a2[1]), "no atomics: results must be bit-reproducible"
    P.set_kernel(api.KERNEL_LDG)
    b1 = P.dual_matvec(xb, xk)
    assert np.array_equal(a1[0], b1[0]) and np.array_equal(a1[1], b1[1]), "both kernels use the same summation order"
    P.close()


# ----------------------------------------------------------------------------- golden fixtures through dyb_propagate
@pytest.mark.parametrize("name", ["prop_N64_dt5e-6", "prop_N128_dt2e-5", "prop_N64_dt5e-4"])
def test_propagate_matches_golden(api, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    N = int(g["N"])
    P = api.Propagator(N)
    P.upload_hprime(g["H_prime"])
    P.set_packets(g["Psi_bra0"], g["Psi_ket0"])               # el + hole in one call, one H' pass per term
    save_tau, traces = P.propagate(float(g["t_init"]), float(g["t_max"]), float(g["tau0"]))
    bra, ket = P.get_packets()
    for p, tag in enumerate(("el", "hl")):
        assert relerr(bra[:, p], g[f"{tag}_bra"]) < REL_TOL
        assert relerr(ket[:, p], g[f"{tag}_ket"]) < REL_TOL
        assert save_tau[p] == float(g[f"{tag}_save_tau"])
        assert traces[p].n_matvec_pairs == int(g[f"{tag}_matvec_pairs"])
        assert traces[p].n_substeps == int(g[f"{tag}_substeps"])
        ev = np.array(events3(traces[p]))
        assert np.array_equal(ev, g[f"{tag}_events"]), "decision trace differs from the oracle's"
        assert np.allclose([e[3] for e in traces[p].events()], g[f"{tag}_event_tau"], rtol=1e-15, atol=0)
    P.close()


@pytest.mark.parametrize("N,dt", [(512, 2e-6), (2048, 5e-7)])
def test_propagate_matches_live_oracle(api, oracle_mod, N, dt):
    w = syn.make_workload(N)
    P = api.Propagator(N)
    Hp = P.form_hprime(w.S, w.h)                              # a2+a3 on the device
    Hp_or = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    assert relerr(Hp, Hp_or) < 1e-11                          # O(cond(S) eps)
    # same H' on both sides for the recursion (conditioning)
    P.upload_hprime(Hp_or)
    P.set_packets(w.Psi_bra, w.Psi_ket)
    tau0 = dt / H_BAR
    save_tau, traces = P.propagate(0.0, dt, tau0)
    bra, ket = P.get_packets()
    for p in range(2):
        b, k, _, st, tr = oracle_mod.propagation(Hp_or, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        assert events3(traces[p]) == events3(tr)
        assert save_tau[p] == st
        assert relerr(bra[:, p], b) < REL_TOL and relerr(ket[:, p], k) < REL_TOL
    P.close()


def test_single_particle_equals_batched(api, golden_dir):
    """el and hole are independent state machines (ElHl_Chebyshev.f:35,185,237): sharing the H' pass must
    not change either result, bit for bit."""
    g = np.load(os.path.join(golden_dir, "prop_N128_dt2e-5.npz"))
    N = int(g["N"])
    P = api.Propagator(N)
    P.upload_hprime(g["H_prime"])
    P.set_packets(g["Psi_bra0"], g["Psi_ket0"])
    P.propagate(float(g["t_init"]), float(g["t_max"]), float(g["tau0"]))
    bra2, ket2 = P.get_packets()
    for p in range(2):
        P.set_packets(g["Psi_bra0"][:, p], g["Psi_ket0"][:, p])
        P.propagate(float(g["t_init"]), float(g["t_max"]), float(g["tau0"]))
        b1, k1 = P.get_packets()
        assert np.array_equal(b1[:, 0], bra2[:, p]) and np.array_equal(k1[:, 0], ket2[:, p])
    P.close()


# ----------------------------------------------------------------------------- the legacy Fortran symbols
def test_legacy_propagationelhl_symbols(api, oracle_mod):
    """propagationelhl_gpucaller_ exactly as ElHl_Chebyshev_GPU.f:269-272 calls it (per particle), and the
    batched el+hole form; compared with the oracle's ElHl_Chebyshev step (ElHl_Chebyshev.f:148-291)."""
    N, dt = 128, 2e-6
    w = syn.make_workload(N)
    st = oracle_mod.ElHlState(w.Psi_bra, w.Psi_ket)
    ref = oracle_mod.elhl_step(st, w.S, w.h, dt)
    tau0 = dt / H_BAR
    out2 = api.legacy_propagationelhl(w.S, w.h, w.Psi_bra, w.Psi_ket, 0.0, dt, tau0)
    assert relerr(out2["H_prime"], ref["H_prime"]) < 1e-11
    assert np.isnan(out2["AO_ket"]).all(), "h_AO_ket must never be touched (Taylor_gpu.cpp:634-736)"
    for p in range(2):
        assert relerr(out2["PSI_bra"][:, p], st.Psi_bra[:, p]) < REL_TOL
        assert relerr(out2["PSI_ket"][:, p], st.Psi_ket[:, p]) < REL_TOL
        # AO_bra is returned UN-conjugated (ElHl_Chebyshev_GPU.f:304 conjugates on the host)
        assert relerr(np.conj(out2["AO_bra"][:, p]), ref["AO_bra"][:, p]) < 1e-9
        assert abs(out2["save_tau"][p] - st.save_tau[p]) <= 1e-15 * st.save_tau[p]
        out1 = api.legacy_propagationelhl(w.S, w.h, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
        assert relerr(out1["PSI_bra"], st.Psi_bra[:, p]) < REL_TOL
        assert relerr(out1["PSI_ket"], st.Psi_ket[:, p]) < REL_TOL
    b, k, sv = api.legacy_propagation(ref["H_prime"], w.Psi_bra[:, 0], w.Psi_ket[:, 0], 0.0, dt, tau0)
    assert relerr(b, st.Psi_bra[:, 0]) < REL_TOL and relerr(k, st.Psi_ket[:, 0]) < REL_TOL
    api.gpu_finalize()


def test_trajectory_populations_match_golden(api, golden_dir):
    """100 nuclear steps with moving nuclei (north_star: populations within 1e-9 over a 100-step trajectory): S,h
    rebuilt on the host each step (reference path), everything else through the library."""
    g = np.load(os.path.join(golden_dir, "traj_N64_dt2e-6_100steps.npz"))
    N, dt, n_steps = int(g["N"]), float(g["dt"]), int(g["n_steps"])
    pos, species = syn.lattice(N // 4, 1234 + N)
    S0, _ = syn.workload_at(pos, species)
    _, Psi_bra, Psi_ket = syn.packets(S0, N)
    frag = syn.fragments(N)
    P = api.Propagator(N)
    P.set_packets(Psi_bra, Psi_ket)
    tau_max = dt / H_BAR
    save_tau = np.zeros(2); t = 0.0
    for step in range(n_steps):
        it = step + 2                                            # Chebyshev_driver.f:106 increments before the call
        S, h = syn.workload_at(syn.perturb_positions(pos, step), species)
        P.form_hprime(S, h, want_hprime=False)
        tau = np.full(2, tau_max) if step == 0 else np.minimum(tau_max, 1.15 * save_tau)   # ElHl_Chebyshev.f:182-184
        t_max = dt * 1 * (it - 1)                                # ElHl_Chebyshev.f:176
        save_tau, traces = P.propagate(t, t_max, tau)
        t = t + dt
        pops = P.populations(frag, 4, t)                         # DUAL_bra = conj(ket), DUAL_ket = bra on the device
        assert np.abs(pops - g["pops"][step]).max() < POP_TOL, f"step {step}"
        assert [tr.n_matvec_pairs for tr in traces] == list(g["pairs"][step])
        assert np.allclose(save_tau, g["save_tau"][step], rtol=1e-14, atol=0)
    bra, ket = P.get_packets()
    assert relerr(bra, g["Psi_bra_final"]) < 1e-9 and relerr(ket, g["Psi_ket_final"]) < 1e-9
    assert relerr(np.conj(P.ao_bra()), g["AO_bra_final"]) < 1e-8
    P.close()


# ----------------------------------------------------------------------------- full-size properties (no oracle needed)
def test_full_size_properties(api):
    """N = 16384 (BASELINE config 3): linearity of the dual product, agreement of the two kernels, bitwise
    repeatability, and consistency <H'^T a, b> == <a, H' b> which ties the bra and ket halves of the pass."""
    import torch
    N = 16384
    P = api.Propagator(N)
    ld = N
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    Hd = torch.empty((N, ld), device="cuda", dtype=torch.float64)     # row j of this tensor = column j of H'
    Hd.normal_(generator=g)
    torch.cuda.synchronize()
    P.upload_hprime_device(Hd.data_ptr(), ld)
    rng = np.random.default_rng(3)
    a = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    b = np.asfortranarray(rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2)))
    ya, yb_ = P.dual_matvec(a, b)                     # ya = H'^T a ; yb_ = H' b
    for p in range(2):
        lhs = np.sum(ya[:, p] * b[:, p]); rhs = np.sum(a[:, p] * yb_[:, p])
        assert abs(lhs - rhs) / abs(lhs) < 1e-11
    # linearity: H(2a + i b) == 2 Ha + i Hb
    y2 = P.dual_matvec(2 * a + 1j * b, 2 * a + 1j * b)
    y_a = P.dual_matvec(a, a); y_b = P.dual_matvec(b, b)
    assert relerr(y2[0], 2 * y_a[0] + 1j * y_b[0]) < 1e-12 and relerr(y2[1], 2 * y_a[1] + 1j * y_b[1]) < 1e-12
    again = P.dual_matvec(a, b)
    assert np.array_equal(again[0], ya) and np.array_equal(again[1], yb_)
    P.set_kernel(api.KERNEL_LDG)
    other = P.dual_matvec(a, b)
    assert np.array_equal(other[0], ya) and np.array_equal(other[1], yb_)
    # against torch fp64 on the same device matrix (plain reference of the same op)
    Hm = Hd[:, :N].t()                                # H' as a torch matrix (view)
    at = torch.tensor(a, device="cuda"); bt = torch.tensor(b, device="cuda")
    ref_k = torch.complex(Hm @ bt.real, Hm @ bt.imag).cpu().numpy()
    ref_b = torch.complex(Hm.t() @ at.real, Hm.t() @ at.imag).cpu().numpy()
    assert relerr(yb_, ref_k) < 1e-12 and relerr(ya, ref_b) < 1e-12
    P.close()


# ----------------------------------------------------------------------------- edge cases of the driver
@pytest.mark.parametrize("N", [5, 130, 259])
def test_propagate_ragged_sizes_random_operator(api, oracle_mod, N):
    """N not a multiple of anything (padding rows/columns, partial TMA boxes, a single partial panel) with a generic
    non-symmetric operator similar to a real-spectrum one (H' = A^-1 D A)."""
    rng = np.random.default_rng(N)
    A = rng.normal(size=(N, N)) + 3.0 * np.eye(N)
    D = np.diag(rng.uniform(-20.0, 60.0, size=N))
    Hp = np.asfortranarray(np.linalg.solve(A, D @ A))
    ket = rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2))
    M = A.T @ A                                            # bra = M ket makes <bra|ket> > 0 and H'^T M = M H' ... not needed exactly
    bra = M @ ket
    nrm = np.sqrt(np.abs(np.einsum("ip,ip->p", np.conj(bra), ket)))
    ket = np.asfortranarray(ket / nrm); bra = np.asfortranarray(bra / nrm)
    dt = 3e-6; tau0 = dt / H_BAR
    P = api.Propagator(N)
    P.upload_hprime(Hp)
    P.set_packets(bra, ket)
    save_tau, traces = P.propagate(0.0, dt, tau0)
    gb, gk = P.get_packets()
    for p in range(2):
        b, k, _, st, tr = oracle_mod.propagation(Hp, bra[:, p], ket[:, p], 0.0, dt, tau0)
        assert events3(traces[p]) == events3(tr) and save_tau[p] == st
        assert relerr(gb[:, p], b) < REL_TOL and relerr(gk[:, p], k) < REL_TOL
    P.close()


def test_propagate_zero_length_slice_still_takes_one_step(api, oracle_mod):
    """t_max == t_init: the reference still runs the first Convergence and advances by tau (Taylor.f:65-81)."""
    N = 64
    w = syn.make_workload(N)
    Hp = oracle_mod.sy_multiply(oracle_mod.sy_invert(w.S), w.h)
    P = api.Propagator(N)
    P.upload_hprime(Hp)
    P.set_packets(w.Psi_bra[:, 0], w.Psi_ket[:, 0])
    tau0 = 1e-4
    save_tau, traces = P.propagate(1.0, 1.0, tau0)
    gb, gk = P.get_packets()
    b, k, _, st, tr = oracle_mod.propagation(Hp, w.Psi_bra[:, 0], w.Psi_ket[:, 0], 1.0, 1.0, tau0)
    assert events3(traces[0]) == events3(tr) and save_tau[0] == st and traces[0].n_substeps == 0
    assert relerr(gb[:, 0], b) < REL_TOL and relerr(gk[:, 0], k) < REL_TOL
    assert not np.array_equal(gk[:, 0], w.Psi_ket[:, 0])
    P.close()


def test_bad_arguments_are_reported(api):
    P = api.Propagator(32)
    with pytest.raises(api.DynemolB200Error) as e:
        P.propagate(0.0, 1e-6, 1e-3)                         # no packets yet
    assert e.value.code == api.EINVAL
    P.set_packets(np.ones((32, 2), complex), np.ones((32, 2), complex))
    with pytest.raises(api.DynemolB200Error) as e:
        P.ao_bra()                                           # needs the factor of S from dyb_form_hprime
    assert "dyb_form_hprime" in str(e.value)
    with pytest.raises(api.DynemolB200Error) as e:
        P.propagate(0.0, 1e-6, 1e-3, mode=api.MODE_CHEBYSHEV)   # no spectral bounds
    assert "spectral bounds" in str(e.value)
    P.close()


def test_li2s_crystal_size_step(api, oracle_mod):
    """BASELINE config 1 (examples/Li2S-crystal: 1728 Li + 864 S, N = 10368 with the 4-orbital model): anti-fluorite
    supercell geometry, S and h from the synthetic EHT recipe, H' formed on the device, one short Taylor step for
    electron and hole against the CPU oracle at full size: identical decisions, 1e-10 wavepackets."""
    import torch
    pos, species = syn.li2s_lattice(6, 6, 6)
    N = 4 * pos.shape[0]
    assert N == 10368
    S_t, h_t, meta = syn.S_h_torch_from_positions(pos, species, torch.device("cuda", 0), zeta=0.45)
    P = api.Propagator(N)
    P.form_hprime_device(S_t.data_ptr(), N, h_t.data_ptr(), N)
    Hp = P.download_hprime()
    S = S_t.cpu().numpy(); h = h_t.cpu().numpy()
    del S_t, h_t
    torch.cuda.empty_cache()
    resid = np.abs(S @ Hp[:, :64] - h[:, :64]).max() / np.abs(h).max()      # S H' = h on a column block
    assert resid < 1e-12
    C, Psi_bra, Psi_ket = syn.packets(np.asfortranarray(S), N)
    P.set_packets(Psi_bra, Psi_ket)
    dt = 4e-7; tau0 = dt / H_BAR
    save_tau, traces = P.propagate(0.0, dt, tau0)
    gb, gk = P.get_packets()
    e = P.quasiparticle_energies()
    for p in range(2):
        b, k, _, st, tr = oracle_mod.propagation(Hp, Psi_bra[:, p], Psi_ket[:, p], 0.0, dt, tau0)
        assert events3(traces[p]) == events3(tr) and save_tau[p] == st
        assert relerr(gb[:, p], b) < REL_TOL and relerr(gk[:, p], k) < REL_TOL
        assert abs(abs(np.vdot(gb[:, p], gk[:, p])) - 1.0) < 1e-7
        assert abs(e[p] - np.vdot(b, Hp @ k)) < 1e-8 * abs(e[p])
    P.close()


def test_legacy_symbol_at_headline_size(api):
    """propagationelhl2_gpucaller_ at N = 16384 (BASELINE config 3) with host S, h exactly as the Fortran caller would
    pass them.  The oracle cannot form S^-1 h at this size in test time, so the check is by properties that pin every
    output: S H' = h, S AO_bra = PSI_bra, charge conservation to the algorithm's 1e-8, untouched AO_ket."""
    import torch
    N = 16384
    S_t, h_t, _ = syn.make_S_h_torch(N, torch.device("cuda", 0))
    S = S_t.cpu().numpy().T; h = h_t.cpu().numpy().T           # symmetric: the transposed views are Fortran-ordered
    del S_t, h_t
    torch.cuda.empty_cache()
    w = 64
    C = np.zeros((N, 2)); C[0:w, 0] = np.random.default_rng(42).normal(size=w); C[w:2 * w, 1] = np.random.default_rng(43).normal(size=w)
    SC = S @ C
    C /= np.sqrt(np.einsum("ip,ip->p", C, SC)); SC = S @ C
    dt = 1e-6
    out = api.legacy_propagationelhl(S, h, SC.astype(np.complex128), C.astype(np.complex128), 0.0, dt, dt / H_BAR)
    Hp = out["H_prime"]
    cols = np.r_[0:16, N - 16:N]
    assert np.abs(S @ Hp[:, cols] - h[:, cols]).max() / np.abs(h).max() < 1e-11                 # H' = S^-1 h
    Sx = lambda z: (S @ z.real) + 1j * (S @ z.imag)            # real matrix x complex vector without upcasting S
    for p in range(2):
        assert np.abs(Sx(out["AO_bra"][:, p]) - out["PSI_bra"][:, p]).max() < 1e-10             # AO_bra = S^-1 PSI_bra
        assert abs(abs(np.vdot(out["PSI_bra"][:, p], out["PSI_ket"][:, p])) - 1.0) < 2e-8       # Taylor.f:104
        assert np.abs(out["PSI_bra"][:, p] - Sx(out["PSI_ket"][:, p])).max() < 1e-7             # bra stays S ket
        assert out["save_tau"][p] > 0
    assert np.isnan(out["AO_ket"]).all()
    api.gpu_finalize()
