"""GPU suite: the slice_Cheb loop mirror (dynemol_b200/driver.py; Chebyshev_driver.f:94-169) -- a trajectory through
the driver against the golden oracle trajectory, and stop/restart through Security_copy.dat against an oracle that is
stopped and restarted the same way (ElHl_Chebyshev.f:377-433: after a restart tau starts again from tau_max)."""
import os

import numpy as np
import pytest

from dynemol_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import api as a
    assert a.device_count() > 0
    return a


def _setup(N):
    pos, species = syn.lattice(N // 4, 1234 + N)
    S0, _ = syn.workload_at(pos, species)
    C, _, _ = syn.packets(S0, N)
    return pos, species, S0, np.asfortranarray(C.astype(np.complex128))


def test_driver_reproduces_golden_trajectory(api, golden_dir):
    from dynemol_b200.driver import SliceChebDriver
    g = np.load(os.path.join(golden_dir, "traj_N64_dt2e-6_100steps.npz"))
    N, dt = int(g["N"]), float(g["dt"])
    pos, species, S0, C = _setup(N)
    drv = SliceChebDriver(N, syn.fragments(N), np.arange(N) // 4, 4, dt)
    p0 = drv.preprocess(S0, C, C)                       # AO_bra = AO_ket = C (real FMO coefficients)
    assert np.allclose(p0[5], 1.0, atol=1e-12)
    for step in range(25):
        S, h = syn.workload_at(syn.perturb_positions(pos, step), species)
        out = drv.step(S, h)
        assert np.abs(out["pops"] - g["pops"][step]).max() < 1e-9, step
        assert [t.n_matvec_pairs for t in out["traces"]] == list(g["pairs"][step])
    assert drv.Net_Charge.shape == (N // 4,) and np.isfinite(drv.Net_Charge).all()   # data_output.f:133-137 (abs per atom)
    drv.close()


def test_stop_and_restart_matches_oracle(api, oracle_mod, tmp_path):
    from dynemol_b200.driver import SliceChebDriver
    N, dt, n1, n2 = 64, 2e-6, 6, 4
    pos, species, S0, C = _setup(N)
    frag = syn.fragments(N)
    geo = lambda step: syn.workload_at(syn.perturb_positions(pos, step), species)

    drv = SliceChebDriver(N, frag, np.arange(N) // 4, 4, dt)
    drv.preprocess(S0, C, C)
    for step in range(n1):
        drv.step(*geo(step))
    path = str(tmp_path / "Security_copy.dat")
    drv.security_copy(path)
    drv.close()

    drv2 = SliceChebDriver(N, frag, np.arange(N) // 4, 4, dt)
    st = drv2.from_restart(path)                        # "mv Security_copy.dat Restart_copy.dat" + restart = .true.
    assert st.it == n1 + 1 and abs(st.t - n1 * dt) < 1e-18
    assert st.frame == n1 + 1                           # the loop value of the last executed step (frame_step = 1: 2 .. n1+1)

    # oracle stopped and restarted the same way: Psi_bra = DUAL_ket, Psi_ket = AO_ket, first_call again
    o = oracle_mod.ElHlState(S0.T @ C, C)
    for step in range(n1):
        last = oracle_mod.elhl_step(o, *geo(step), dt)
    assert np.abs(st.DUAL_ket - last["DUAL_ket"]).max() < 1e-9 and np.abs(st.AO_ket - last["AO_ket"]).max() < 1e-9
    o2 = oracle_mod.ElHlState(st.DUAL_ket, st.AO_ket)
    o2.t, o2.it, o2.first_call = st.t, st.it, True
    for step in range(n1, n1 + n2):
        out = drv2.step(*geo(step))
        ref = oracle_mod.elhl_step(o2, *geo(step), dt)
        pops_ref = oracle_mod.populations(frag, ref["DUAL_bra"], ref["DUAL_ket"], ref["t"], 4)
        assert np.abs(out["pops"] - pops_ref).max() < 1e-9
        assert [t.n_matvec_pairs for t in out["traces"]] == [t.n_matvec_pairs for t in ref["traces"]]
        erg_ref = oracle_mod.quasiparticle_energies(ref["AO_bra"], ref["AO_ket"], geo(step)[1])
        assert np.abs(out["erg"] - erg_ref).max() < 1e-8 * np.abs(erg_ref).max()
    drv2.close()


def test_frame_numbers_follow_the_reference_loop(api, tmp_path):
    """Chebyshev_driver.f:100-165: do frame = frame_step+1, frame_final, frame_step, and Security_Copy stores the loop
    value; after a restart the loop starts at frame_restart + 1.  With frame_step = 3: 4, 7, 10 | restart at 10 -> 11, 14."""
    from dynemol_b200.driver import SliceChebDriver
    from dynemol_b200.restart import read_restart_copy
    N, dt, fs = 64, 1e-6, 3
    pos, species, S0, C = _setup(N)
    frag = syn.fragments(N)
    S, h = syn.workload_at(pos, species)
    drv = SliceChebDriver(N, frag, np.arange(N) // 4, 4, dt, frame_step=fs)
    drv.preprocess(S0, C, C)
    seen = []
    for _ in range(3):
        drv.step(S, h); seen.append(drv.frame)
    assert seen == [4, 7, 10]
    path = str(tmp_path / "Security_copy.dat")
    drv.security_copy(path)
    drv.close()
    assert read_restart_copy(path).frame == 10
    drv2 = SliceChebDriver(N, frag, np.arange(N) // 4, 4, dt, frame_step=fs)
    drv2.from_restart(path)
    drv2.step(S, h); a = drv2.frame
    drv2.step(S, h); b = drv2.frame
    assert (a, b) == (11, 14)
    drv2.close()
