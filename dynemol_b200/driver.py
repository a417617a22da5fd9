"""Host-side mirror of the reference's `slice_Cheb` time loop for the propagator path:
Chebyshev_driver.f:94-169 (loop, `it` bookkeeping, security copies) around ElHl_Chebyshev.f:148-291 (one nuclear
step) and preprocess_ElHl_Chebyshev / preprocess_from_restart (ElHl_Chebyshev.f:50-143, 377-433).

What stays on the reference path (and is therefore an INPUT here): the geometry, the overlap S and the Hueckel
matrix h of every frame (overlap_D.f, hamiltonians.f), MM dynamics, environment fields.  Everything from S, h to the
fragment populations runs through libdynemol_b200.so.
"""
from __future__ import annotations

import numpy as np

from . import api
from .restart import ChebState, read_restart_copy, write_security_copy

H_BAR = api.H_BAR


class SliceChebDriver:
    def __init__(self, N: int, fragment, atom, n_frag: int, delta_t: float, frame_step: int = 1, t_i: float = 0.0,
                 mode: int = api.MODE_TAYLOR, device: int = 0, eh_tag=("el", "hl")):
        self.N, self.n_frag, self.delta_t, self.frame_step, self.mode = N, n_frag, delta_t, frame_step, mode
        self.fragment = np.ascontiguousarray(fragment, dtype=np.int32)
        self.atom = np.ascontiguousarray(atom, dtype=np.int64)
        self.n_atoms = int(self.atom.max()) + 1
        self.eh_tag = list(eh_tag)
        self.P = api.Propagator(N, device=device)
        self.t, self.it = t_i, 1                      # Chebyshev_driver.f:80-83: it = 1, t = t_i
        self.first_call = True                        # ElHl_Chebyshev.f:34 first_call_
        self.save_tau = np.zeros(2)                   # ElHl_Chebyshev.f:35
        self.AO_bra = self.AO_ket = self.DUAL_bra = self.DUAL_ket = None
        self.Net_Charge = np.zeros(self.n_atoms)
        # Chebyshev_driver.f:100-102: do frame = frame_init, frame_final, frame_step with frame_init = frame_step + 1
        # (frame_restart + 1 after a restart); `frame` is the loop value of the step in progress / last executed, which
        # is what Security_Copy stores (:165)
        self.frame = 1
        self._next_frame = frame_step + 1

    # -------------------------------------------------------------------------------------------------- start
    def preprocess(self, S, AO_bra, AO_ket):
        """preprocess_ElHl_Chebyshev (ElHl_Chebyshev.f:50-143): DUAL_bra = AO_bra, DUAL_ket = S AO_ket,
        Psi_bra = AO_bra^T S, Psi_ket = AO_ket; returns the populations at t_i."""
        S = np.asarray(S)
        self.AO_bra = np.asfortranarray(AO_bra, dtype=np.complex128); self.AO_ket = np.asfortranarray(AO_ket, dtype=np.complex128)
        self.DUAL_bra = self.AO_bra.copy(order="F")
        self.DUAL_ket = np.asfortranarray(S @ self.AO_ket)
        Psi_bra = np.asfortranarray(S.T @ self.AO_bra)
        self.P.set_packets(Psi_bra, self.AO_ket)
        return self._populations_host(self.DUAL_bra, self.DUAL_ket, self.t)

    def from_restart(self, path: str):
        """Restart_stuff (Chebyshev_driver.f:331-362) + preprocess_from_restart (ElHl_Chebyshev.f:377-433):
        Psi_bra = DUAL_ket, Psi_ket = AO_ket; first_call_ is true again, so the next step starts from tau_max."""
        st = read_restart_copy(path)
        self.frame, self.it, self.t, self.eh_tag = st.frame, st.it, st.t, st.eh_tag
        self._next_frame = st.frame + 1                                       # Chebyshev_driver.f:100: frame_init = frame_restart + 1
        self.DUAL_bra, self.DUAL_ket, self.AO_bra, self.AO_ket = st.DUAL_bra, st.DUAL_ket, st.AO_bra, st.AO_ket
        self.Net_Charge = st.Net_Charge
        self.P.set_packets(self.DUAL_ket, self.AO_ket)
        self.first_call = True
        return st

    # -------------------------------------------------------------------------------------------------- one frame
    def step(self, S, h, want_hprime: bool = False):
        """One pass of the loop body Chebyshev_driver.f:102-108 -> ElHl_Chebyshev (ElHl_Chebyshev.f:148-291)."""
        self.frame = self._next_frame; self._next_frame += self.frame_step    # Chebyshev_driver.f:102
        self.it += 1                                                          # Chebyshev_driver.f:106
        t_init = self.t
        t_max = self.delta_t * self.frame_step * (self.it - 1)                # ElHl_Chebyshev.f:176
        tau_max = self.delta_t / H_BAR                                        # :178
        tau = np.full(2, tau_max) if self.first_call else np.minimum(tau_max, 1.15 * self.save_tau)   # :182-184
        Hp = self.P.form_hprime(S, h, want_hprime=want_hprime)                # :206-210 on the device
        if self.mode in (api.MODE_CHEBYSHEV, api.MODE_CHEBYSHEV_FULL):
            self.P.estimate_spectral_bounds(24, 0.05)
        self.save_tau[: self.P.n_part], traces = self.P.propagate(t_init, t_max, tau[: self.P.n_part], mode=self.mode)   # :228,253
        self.t = t_init + self.delta_t * self.frame_step                      # :266
        bra, ket = self.P.get_packets()
        self.DUAL_bra = np.conj(ket); self.DUAL_ket = bra                     # :270-271
        self.AO_bra = np.conj(self.P.ao_bra()); self.AO_ket = ket             # :274-276
        erg = self.P.quasiparticle_energies()                                 # :278
        pops = self.P.populations(self.fragment, self.n_frag, self.t)         # :281 on the device
        self._net_charge(self.DUAL_bra, self.DUAL_ket)
        self.first_call = False
        return dict(pops=pops, erg=erg, traces=traces, t=self.t, H_prime=Hp)

    # -------------------------------------------------------------------------------------------------- helpers
    def _net_charge(self, bra, ket):
        """data_output.f:133-137: Net_Charge(atom) = sum_n ChargeSign(n) * | sum_{i in atom} bra(i,n) ket(i,n) |."""
        sign = (-1.0, 1.0)
        nc = np.zeros(self.n_atoms)
        for n in range(bra.shape[1]):
            prod = bra[:, n] * ket[:, n]
            per_atom = np.zeros(self.n_atoms, dtype=np.complex128)
            np.add.at(per_atom, self.atom, prod)
            nc += sign[n] * np.abs(per_atom)
        self.Net_Charge = nc

    def _populations_host(self, bra, ket, t):
        out = np.zeros((self.n_frag + 2, bra.shape[1]))
        for n in range(bra.shape[1]):
            prod = bra[:, n] * ket[:, n]
            out[0, n] = t
            for f in range(self.n_frag):
                out[1 + f, n] = prod[self.fragment == f].sum().real
            out[1 + self.n_frag, n] = prod.sum().real
        self._net_charge(bra, ket)
        return out

    def security_copy(self, path: str):
        """Security_Copy_Cheb (backup.f:329-394), called every step_security frames (Chebyshev_driver.f:165)."""
        write_security_copy(path, ChebState(self.frame, self.it, self.t, self.eh_tag, self.DUAL_bra, self.DUAL_ket,
                                            self.AO_bra, self.AO_ket, self.Net_Charge))

    def close(self):
        self.P.close()
