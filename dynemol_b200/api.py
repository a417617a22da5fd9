"""Host-side mirror of the reference interface for the propagator path, over the C ABI
(include/dynemol_b200.h) with ctypes.  No torch types cross this boundary: numpy host
buffers in, numpy host buffers out, exactly like the Fortran caller's arrays.

  * `legacy_*` wrappers call the Fortran-mangled symbols by reference, the way ifort/ifx
    would (ElHl_Chebyshev_GPU.f:269-272), so parity tests read like the reference's call site.
  * `Propagator` wraps the native dyb_* handle API (device-resident H' and wavepackets).

The CUDA library is mandatory: importing this module without libdynemol_b200.so raises, and
every compute call fails loudly without a GPU (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DYNEMOL_B200_LIB") or os.path.join(_HERE, "lib", "libdynemol_b200.so")   # override: tuning builds

H_BAR = 6.58264e-4          # eV*ps (constants_m.f:23)
MODE_TAYLOR, MODE_CHEBYSHEV, MODE_TAYLOR_REFGPU, MODE_CHEBYSHEV_REFGPU, MODE_CHEBYSHEV_FULL = 0, 1, 2, 3, 4
KERNEL_AUTO, KERNEL_TMA, KERNEL_LDG = 0, 1, 2
MAX_EVENTS = 256

OK, ENODEV, ECUDA, EINVAL, ENOMEM, ESINGULAR = 0, -1, -2, -3, -4, -5


class DynemolB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dynemol_b200 error {code}: {msg}")
        self.code = code


class Trace(C.Structure):
    """Mirror of dyb_trace."""
    _fields_ = [
        ("n_convergence_calls", C.c_int32), ("n_substeps", C.c_int32), ("n_matvec_pairs", C.c_int32),
        ("n_rescale", C.c_int32), ("n_first_shrink", C.c_int32), ("last_k_ref", C.c_int32),
        ("n_events", C.c_int32),
        ("ev_kind", C.c_int32 * MAX_EVENTS), ("ev_k", C.c_int32 * MAX_EVENTS), ("ev_ok", C.c_int32 * MAX_EVENTS),
        ("ev_tau", C.c_double * MAX_EVENTS),
        ("norm_ref", C.c_double), ("final_tau", C.c_double),
    ]

    def events(self):
        n = min(self.n_events, MAX_EVENTS)
        return [(self.ev_kind[i], self.ev_k[i], self.ev_ok[i], self.ev_tau[i]) for i in range(n)]

    def summary(self):
        return dict(convergence_calls=self.n_convergence_calls, substeps=self.n_substeps,
                    matvec_pairs=self.n_matvec_pairs, rescale=self.n_rescale,
                    first_shrink=self.n_first_shrink, k_ref=self.last_k_ref,
                    norm_ref=self.norm_ref, final_tau=self.final_tau)


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dynemol_b200.build` "
            "(nvcc, sm_100a).  dynemol_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.dyb_last_error.restype = C.c_char_p
    lib.dyb_version.restype = C.c_char_p
    lib.dyb_launch_count.restype = C.c_int64
    lib.dyb_launch_count.argtypes = [C.c_void_p]
    lib.nakedbessel_.restype = C.c_double
    lib.dyb_team_last_error.restype = C.c_char_p
    lib.dyb_team_passes_last.restype = C.c_int64
    lib.dyb_team_passes_last.argtypes = [C.c_void_p]
    lib.dyb_unwrap_pin_bytes.restype = C.c_int64
    lib.dyb_legacy_passes_last.restype = C.c_int64
    return lib


lib = _load()

# every symbol include/dynemol_b200.h declares (checked by the CPU test-suite)
DECLARED_SYMBOLS = [
    "propagationelhl_gpucaller_", "propagationelhl2_gpucaller_", "propagation_gpucaller_", "nakedbessel_", "ehrenfestkernel_gpu_", "ehrenfestkernel2_gpu_",
    "gpu_init_", "gpu_finalize_", "gpu_pin_", "gpu_unpin_", "xpu_syinvert_", "xpu_dsymm_", "xpu_dzgemv_",
    "dyb_team_create", "dyb_team_destroy", "dyb_team_size", "dyb_team_form_hprime", "dyb_team_wait_outputs", "dyb_team_upload_hprime",
    "dyb_team_set_packets", "dyb_team_get_packets", "dyb_team_set_spectral_bounds", "dyb_team_estimate_spectral_bounds",
    "dyb_team_propagate", "dyb_team_ao_bra", "dyb_team_populations", "dyb_team_quasiparticle_energies", "dyb_team_run_terms",
    "dyb_team_passes_last", "dyb_team_last_error", "dyb_unwrap_pin_bytes", "dyb_legacy_passes_last",
    "dyb_form_hprime_async", "dyb_wait_outputs", "dyb_factor_overlap", "dyb_factor_device", "dyb_upload_column_block",
    "dyb_solve_column_block", "dyb_column_block_device", "dyb_take_rows_from_column_blocks", "dyb_download_hprime_rows_device", "dyb_comm_p2p_open_local",
    "dyb_last_error", "dyb_version", "dyb_device_count", "dyb_plan", "dyb_resident_plan", "dyb_mid_plan", "dyb_steady_schedule", "dyb_series_coefficients", "dyb_create", "dyb_destroy", "dyb_set_kernel", "dyb_set_series_kernel",
    "dyb_get_info", "dyb_upload_hprime", "dyb_upload_hprime_device", "dyb_upload_hprime_rows_device", "dyb_hprime_device", "dyb_form_hprime", "dyb_form_hprime_device", "dyb_form_hprime_from_overlap",
    "dyb_download_hprime", "dyb_set_packets", "dyb_get_packets", "dyb_propagate", "dyb_ao_bra",
    "dyb_populations", "dyb_run_terms", "dyb_dual_matvec", "dyb_sync", "dyb_launch_count",
    "dyb_comm_unique_id", "dyb_comm_init", "dyb_comm_p2p_handle", "dyb_comm_p2p_open", "dyb_comm_p2p_enable", "dyb_set_spectral_bounds", "dyb_get_spectral_bounds", "dyb_estimate_spectral_bounds",
    "dyb_quasiparticle_energies", "dyb_ehrenfest_kernel", "dyb_ehrenfest_kernel2",
]


def _check(rc: int):
    if rc != 0:
        raise DynemolB200Error(rc, lib.dyb_last_error().decode())


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _fz(a, copy=True) -> np.ndarray:
    return np.array(a, dtype=np.complex128, order="F", copy=copy)


def _fd(a) -> np.ndarray:
    return np.asfortranarray(a, dtype=np.float64)


def _prefer_torch_nccl():
    """The library binds NCCL with dlopen("libnccl.so.2"): whichever copy the process loaded first wins, and torch's
    libtorch_cuda.so needs ITS bundled (newer) copy.  If torch is installed but not imported yet, import it before the
    first NCCL use so that a later `import torch` in the same process still works (plumbing only)."""
    import sys
    if "torch" not in sys.modules:
        try:
            import torch  # noqa: F401
        except Exception:
            pass


def comm_unique_id() -> bytes:
    _prefer_torch_nccl()
    buf = C.create_string_buffer(128)
    _check(lib.dyb_comm_unique_id(buf))
    return buf.raw


def plan(N: int, n_rows: int | None = None, sm_count: int = 148) -> dict:
    """Launch plan of the dual product (host arithmetic only, works without a GPU)."""
    n_rows = N if n_rows is None else n_rows
    out = (C.c_int64 * 8)()
    _check(lib.dyb_plan(C.c_int(N), C.c_int(n_rows), C.c_int(sm_count), out, None, None))
    d = dict(zip(["panels", "tiles_per_panel", "tiles", "grid", "segments", "tile_cols", "panel_rows", "padded_cols"], [int(v) for v in out]))
    seg_base = (C.c_int32 * d["grid"])(); pseg = (C.c_int32 * (d["panels"] + 1))()
    _check(lib.dyb_plan(C.c_int(N), C.c_int(n_rows), C.c_int(sm_count), out, seg_base, pseg))
    d["seg_base"] = list(seg_base); d["pseg_start"] = list(pseg)
    return d


def resident_plan(N: int, sm_count: int = 148, smem_optin: int = 232448) -> dict:
    """Blocking of the shared-memory-resident series kernel (host arithmetic only, works without a GPU)."""
    out = (C.c_int64 * 6)()
    _check(lib.dyb_resident_plan(C.c_int(N), C.c_int(sm_count), C.c_int64(smem_optin), out))
    return dict(zip(["grid_side", "block", "smem_stride", "smem_bytes", "threads", "fits"], [int(v) for v in out]))


def mid_plan(N: int, sm_count: int = 148, smem_optin: int = 232448) -> dict:
    """Blocking of the streamed one-launch series kernel for mid-size operators (csrc/mid.cuh); host arithmetic only."""
    out = (C.c_int64 * 12)()
    _check(lib.dyb_mid_plan(C.c_int(N), C.c_int(sm_count), C.c_int64(smem_optin), out))
    d = dict(zip(["block_rows", "tile_cols", "grid_rows", "grid_cols", "block_cols", "tiles_per_term", "stages",
                  "owned", "owners", "collect_words", "smem_bytes", "fits"], [int(v) for v in out]))
    d["table16"] = 1 if d["collect_words"] < 0 else 0
    d["collect_words"] = abs(d["collect_words"])
    return d


def steady_schedule(t: float, t_max: float, tau: float, max_sub: int = 4096) -> np.ndarray:
    """tau of every remaining steady sub-step (Taylor.f:81-126) if all norm tests pass; host arithmetic only."""
    out = np.zeros(max_sub)
    lib.dyb_steady_schedule.restype = C.c_int
    n = lib.dyb_steady_schedule(C.c_double(t), C.c_double(t_max), C.c_double(tau), C.c_int(max_sub), _p(out))
    if n < 0:
        raise ValueError("dyb_steady_schedule: bad argument")
    return out[:n].copy()


def series_coefficients(mode: int, tau: float, ebar: float = 0.0, de: float = 1.0):
    """(C[25], k_max) the library uses for this tau (host arithmetic only)."""
    out = np.zeros(25, dtype=np.complex128); km = C.c_int(0)
    _check(lib.dyb_series_coefficients(C.c_int(mode), C.c_double(tau), C.c_double(ebar), C.c_double(de), _p(out), C.byref(km)))
    return out, km.value


def device_count() -> int:
    return int(lib.dyb_device_count())


def unwrap_pin_bytes(size_bytes_i32: int) -> int:
    """The byte count gpu_pin_ recovers from the wrapped 32-bit `n*n*8` the Fortran caller passes (host arithmetic)."""
    return int(lib.dyb_unwrap_pin_bytes(C.c_int(np.int64(size_bytes_i32).astype(np.int32))))


def _tcheck(rc: int):
    if rc != 0:
        raise DynemolB200Error(rc, lib.dyb_team_last_error().decode())


class Team:
    """Single-process multi-GPU: `n_dev` row-sharded contexts behind one caller (include/dynemol_b200.h, dyb_team_*).
    Same host-buffer interface as Propagator; H' = S^-1 h is formed on the first device and scattered over NVLink."""

    def __init__(self, N: int, n_dev: int, devices=None):
        self.N, self.n_dev = int(N), int(n_dev)
        self._h = C.c_void_p()
        if n_dev > 1:
            _prefer_torch_nccl()
        devs = None if devices is None else (C.c_int * n_dev)(*devices)
        _tcheck(lib.dyb_team_create(C.byref(self._h), C.c_int(n_dev), devs, C.c_int(N)))
        self.n_part = 0

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = C.c_void_p()
            try:
                lib.dyb_team_destroy(h)
            except Exception:
                pass

    __del__ = close

    def form_hprime(self, S, h, want_hprime: bool = True):
        S = _fd(S); h = _fd(h)
        out = np.empty((self.N, self.N), dtype=np.float64, order="F") if want_hprime else None
        _tcheck(lib.dyb_team_form_hprime(self._h, _p(S), _p(h), _p(out) if want_hprime else None))
        _tcheck(lib.dyb_team_wait_outputs(self._h))
        return out

    def upload_hprime(self, H):
        H = _fd(H)
        _tcheck(lib.dyb_team_upload_hprime(self._h, _p(H), C.c_int64(self.N)))

    def set_packets(self, bra, ket):
        bra = _fz(bra); ket = _fz(ket)
        if bra.ndim == 1:
            bra = np.asfortranarray(bra[:, None]); ket = np.asfortranarray(ket[:, None])
        self.n_part = bra.shape[1]
        _tcheck(lib.dyb_team_set_packets(self._h, C.c_int(self.n_part), _p(bra), _p(ket)))

    def get_packets(self):
        bra = np.empty((self.N, self.n_part), dtype=np.complex128, order="F"); ket = np.empty_like(bra, order="F")
        _tcheck(lib.dyb_team_get_packets(self._h, C.c_int(self.n_part), _p(bra), _p(ket)))
        return bra, ket

    def set_spectral_bounds(self, emin, emax):
        _tcheck(lib.dyb_team_set_spectral_bounds(self._h, C.c_double(emin), C.c_double(emax)))

    def estimate_spectral_bounds(self, n_iter: int = 40, margin: float = 0.05):
        lo = C.c_double(); hi = C.c_double()
        _tcheck(lib.dyb_team_estimate_spectral_bounds(self._h, C.c_int(n_iter), C.c_double(margin), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def propagate(self, t_init, t_max, tau, mode: int = MODE_TAYLOR):
        tau2 = np.zeros(2); tau2[: self.n_part] = np.broadcast_to(np.asarray(tau, dtype=np.float64), (self.n_part,))
        save = np.zeros(2)
        traces = (Trace * 2)()
        _tcheck(lib.dyb_team_propagate(self._h, C.c_int(mode), C.c_double(t_init), C.c_double(t_max), _p(tau2), _p(save), traces))
        return save[: self.n_part].copy(), [traces[i] for i in range(self.n_part)]

    def ao_bra(self):
        out = np.empty((self.N, self.n_part), dtype=np.complex128, order="F")
        _tcheck(lib.dyb_team_ao_bra(self._h, C.c_int(self.n_part), _p(out)))
        return out

    def populations(self, fragment, n_frag: int, t: float):
        frag = np.ascontiguousarray(fragment, dtype=np.int32)
        out = np.zeros((n_frag + 2, self.n_part), dtype=np.float64, order="F")
        _tcheck(lib.dyb_team_populations(self._h, C.c_int(self.n_part), C.c_int(n_frag), _p(frag), C.c_double(t), _p(out)))
        return out

    def quasiparticle_energies(self):
        out = np.zeros(4)
        _tcheck(lib.dyb_team_quasiparticle_energies(self._h, C.c_int(self.n_part), _p(out)))
        return (out[0::2] + 1j * out[1::2])[: self.n_part]

    def run_terms(self, tau: float, n_terms: int):
        ms = C.c_float(0.0)
        _tcheck(lib.dyb_team_run_terms(self._h, C.c_double(tau), C.c_int(n_terms), C.byref(ms)))
        return ms.value

    def passes_last(self) -> int:
        return int(lib.dyb_team_passes_last(self._h))


class Propagator:
    """Native handle API: one GPU, one basis size, H' and the wavepackets resident in HBM."""

    def __init__(self, N: int, device: int = 0, row0: int = 0, n_rows: int | None = None, kernel: int = KERNEL_AUTO):
        self.N = int(N)
        self._h = C.c_void_p()
        _check(lib.dyb_create(C.byref(self._h), C.c_int(device), C.c_int(N), C.c_int(row0),
                              C.c_int(N if n_rows is None else n_rows)))
        if kernel != KERNEL_AUTO:
            self.set_kernel(kernel)
        self.n_part = 0

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = C.c_void_p()
            try:
                lib.dyb_destroy(h)
            except Exception:          # interpreter shutdown: the library may already be gone
                pass

    __del__ = close

    def set_kernel(self, kernel: int):
        _check(lib.dyb_set_kernel(self._h, C.c_int(kernel)))

    def set_series_kernel(self, kind):
        """kind: 'auto' | 'term' | 'resident' | 'mid' (include/dynemol_b200.h: DYB_SERIES_*)."""
        code = {"auto": 0, "term": 1, "resident": 3, "mid": 5}[kind] if isinstance(kind, str) else int(kind)
        _check(lib.dyb_set_series_kernel(self._h, C.c_int(code)))

    def info(self) -> dict:
        buf = (C.c_int64 * 16)()
        _check(lib.dyb_get_info(self._h, buf))
        keys = ["N", "ld", "n_rows", "grid", "tiles", "segments", "sm_count", "smem_bytes", "variant", "panels", "tiles_per_panel", "passes_last", "p2p", "series_kernel", "resident_grid_side", "resident_block"]
        return {k: int(buf[i]) for i, k in enumerate(keys)}

    # ---- operator
    def upload_hprime(self, H):
        H = _fd(H)
        assert H.shape == (self.N, self.N)
        _check(lib.dyb_upload_hprime(self._h, _p(H), C.c_int64(self.N)))

    def upload_hprime_device(self, d_ptr: int, lda: int):
        """H' from a device buffer (column-major, leading dimension lda), e.g. a torch tensor's data_ptr()."""
        _check(lib.dyb_upload_hprime_device(self._h, C.c_void_p(d_ptr), C.c_int64(lda)))

    def upload_hprime_rows_device(self, d_ptr: int, lda: int, local_row0: int, n_rows: int):
        """A block of owned rows of H' from a device buffer holding only those rows (n_rows x N, column-major)."""
        _check(lib.dyb_upload_hprime_rows_device(self._h, C.c_void_p(d_ptr), C.c_int64(lda), C.c_int(local_row0), C.c_int(n_rows)))

    def download_hprime_rows_device(self, d_ptr: int, ldd: int, local_row0: int, n_rows: int):
        """Copy owned rows local_row0.. of the resident H' (all N columns) into a column-major device buffer (n_rows x N)."""
        _check(lib.dyb_download_hprime_rows_device(self._h, C.c_void_p(d_ptr), C.c_int64(ldd), C.c_int(local_row0), C.c_int(n_rows)))

    def hprime_device(self):
        ptr = C.c_void_p(); ld = C.c_int64()
        _check(lib.dyb_hprime_device(self._h, C.byref(ptr), C.byref(ld)))
        return ptr.value, ld.value

    def form_hprime(self, S, h, want_hprime: bool = True):
        """a2+a3: H' = S^-1 h on the device (ElHl_Chebyshev.f:206-210)."""
        S = _fd(S); h = _fd(h)
        out = np.empty((self.N, self.N), dtype=np.float64, order="F") if want_hprime else None
        _check(lib.dyb_form_hprime(self._h, _p(S), _p(h), _p(out) if want_hprime else None))
        return out

    def form_hprime_from_overlap(self, S, IP, k_WH, V_shift, want_hprime: bool = True):
        """Build_Huckel on the device (h = X o S) followed by H' = S^-1 h; only S is uploaded."""
        S = _fd(S); IP = np.ascontiguousarray(IP, dtype=np.float64); k_WH = np.ascontiguousarray(k_WH, dtype=np.float64)
        V_shift = np.ascontiguousarray(V_shift, dtype=np.float64)
        out = np.empty((self.N, self.N), dtype=np.float64, order="F") if want_hprime else None
        _check(lib.dyb_form_hprime_from_overlap(self._h, _p(S), _p(IP), _p(k_WH), _p(V_shift), _p(out) if want_hprime else None))
        return out

    def form_hprime_device(self, d_S: int, lds: int, d_h: int, ldh: int):
        _check(lib.dyb_form_hprime_device(self._h, C.c_void_p(d_S), C.c_int64(lds), C.c_void_p(d_h), C.c_int64(ldh)))

    def download_hprime(self):
        out = np.empty((self.N, self.N), dtype=np.float64, order="F")
        _check(lib.dyb_download_hprime(self._h, _p(out), C.c_int64(self.N)))
        return out

    # ---- packets
    def set_packets(self, bra, ket):
        bra = _fz(bra); ket = _fz(ket)
        if bra.ndim == 1:
            bra = np.asfortranarray(bra[:, None]); ket = np.asfortranarray(ket[:, None])
        assert bra.shape == ket.shape and bra.shape[0] == self.N and bra.shape[1] in (1, 2)
        self.n_part = bra.shape[1]
        _check(lib.dyb_set_packets(self._h, C.c_int(self.n_part), _p(bra), _p(ket)))

    def get_packets(self):
        bra = np.empty((self.N, self.n_part), dtype=np.complex128, order="F"); ket = np.empty_like(bra, order="F")
        _check(lib.dyb_get_packets(self._h, C.c_int(self.n_part), _p(bra), _p(ket)))
        return bra, ket

    # ---- propagation
    def propagate(self, t_init, t_max, tau, mode: int = MODE_TAYLOR):
        """a4/a5.  Returns (save_tau[n_part], [Trace...])."""
        tau = np.ascontiguousarray(np.broadcast_to(np.asarray(tau, dtype=np.float64), (self.n_part,)))
        tau2 = np.zeros(2); tau2[: self.n_part] = tau
        save = np.zeros(2)
        traces = (Trace * 2)()
        _check(lib.dyb_propagate(self._h, C.c_int(mode), C.c_double(t_init), C.c_double(t_max), _p(tau2), _p(save), traces))
        return save[: self.n_part].copy(), [traces[i] for i in range(self.n_part)]

    def set_spectral_bounds(self, emin: float, emax: float):
        _check(lib.dyb_set_spectral_bounds(self._h, C.c_double(emin), C.c_double(emax)))

    def estimate_spectral_bounds(self, n_iter: int = 40, margin: float = 0.05):
        """Lanczos (S inner product) from the current packets; returns and stores (emin, emax)."""
        lo = C.c_double(); hi = C.c_double()
        _check(lib.dyb_estimate_spectral_bounds(self._h, C.c_int(n_iter), C.c_double(margin), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def run_terms(self, tau: float, n_terms: int, per_kernel: bool = False):
        """n_terms el+hole series terms, no host decisions.  Returns (elapsed_ms, matvec_kernel_ms or None)."""
        el = C.c_float(0.0); km = C.c_float(0.0)
        _check(lib.dyb_run_terms(self._h, C.c_double(tau), C.c_int(n_terms), C.byref(el), C.byref(km) if per_kernel else None))
        return el.value, (km.value if per_kernel else None)

    def dual_matvec(self, xb, xk):
        """(H'^T xb, H' xk) in one pass over H' -- kernel-level parity entry."""
        xb = _fz(xb); xk = _fz(xk)
        if xb.ndim == 1:
            xb = np.asfortranarray(xb[:, None]); xk = np.asfortranarray(xk[:, None])
        n_part = xb.shape[1]
        yb = np.empty_like(xb, order="F"); yk = np.empty_like(xk, order="F")
        _check(lib.dyb_dual_matvec(self._h, C.c_int(n_part), _p(xb), _p(xk), _p(yb), _p(yk)))
        return yb, yk

    # ---- post-step
    def ao_bra(self):
        out = np.empty((self.N, self.n_part), dtype=np.complex128, order="F")
        _check(lib.dyb_ao_bra(self._h, C.c_int(self.n_part), _p(out)))
        return out

    def populations(self, fragment, n_frag: int, t: float):
        frag = np.ascontiguousarray(fragment, dtype=np.int32)
        out = np.zeros((n_frag + 2, self.n_part), dtype=np.float64, order="F")
        _check(lib.dyb_populations(self._h, C.c_int(self.n_part), C.c_int(n_frag), _p(frag), C.c_double(t), _p(out)))
        return out

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        assert len(unique_id) == 128
        _check(lib.dyb_comm_init(self._h, C.c_int(rank), C.c_int(world), C.c_char_p(unique_id)))

    def quasiparticle_energies(self):
        """ElHl_Chebyshev.f:329-371 on the device: complex energy per particle."""
        out = np.zeros(4)
        _check(lib.dyb_quasiparticle_energies(self._h, C.c_int(self.n_part), _p(out)))
        return (out[0::2] + 1j * out[1::2])[: self.n_part]

    def ehrenfest_kernel(self, A, X):
        """K = X o A - H' A with the resident H' (diabatic-Ehren.f:115-119)."""
        A = _fd(A); X = _fd(X)
        K = np.empty((self.N, self.N), dtype=np.float64, order="F")
        _check(lib.dyb_ehrenfest_kernel(self._h, _p(A), _p(X), _p(K)))
        return K

    def comm_p2p_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib.dyb_comm_p2p_handle(self._h, buf))
        return buf.raw

    def comm_p2p_open(self, handles: bytes):
        _check(lib.dyb_comm_p2p_open(self._h, C.c_char_p(handles)))

    def comm_p2p_enable(self, on: bool):
        _check(lib.dyb_comm_p2p_enable(self._h, C.c_int(1 if on else 0)))

    def sync(self):
        _check(lib.dyb_sync(self._h))

    def launch_count(self) -> int:
        return int(lib.dyb_launch_count(self._h))


# --------------------------------------------------------------------------- legacy Fortran symbols, called by reference
def _ref(x, ctype):
    return C.byref(ctype(x))


def legacy_propagationelhl(S, h, PSI_bra, PSI_ket, t_init, t_max, tau, batched: bool | None = None, copy_inputs: bool = True, out_H=None):
    """call PropagationElHl_gpucaller(N, S, h0, H_prime, AO_bra(:,p), AO_ket(:,p), Psi_t_bra(:,p), Psi_t_ket(:,p),
    t_init, t_max, tau, save_tau)  -- ElHl_Chebyshev_GPU.f:269-272.  PSI_* of shape (N,) use the per-particle
    symbol, shape (N,2) the batched el+hole symbol.  Returns dict(H_prime, AO_bra, PSI_bra, PSI_ket, save_tau)."""
    if copy_inputs:
        S = np.array(S, dtype=np.float64, order="F", copy=True); h = np.array(h, dtype=np.float64, order="F", copy=True)
    else:                      # S and h are read-only for the callee (const double* in the ABI)
        S = _fd(S); h = _fd(h)
    N = S.shape[0]
    PSI_bra = _fz(PSI_bra); PSI_ket = _fz(PSI_ket)
    two = PSI_bra.ndim == 2 and PSI_bra.shape[1] == 2
    if batched is None:
        batched = two
    # out_H: the caller's persistent (N, N) Fortran-ordered H_prime array (the Fortran caller allocates and pins it once,
    # ElHl_Chebyshev_GPU.f:109-111); default: a fresh one per call
    Hp = np.zeros((N, N), dtype=np.float64, order="F") if out_H is None else out_H
    assert Hp.shape == (N, N) and Hp.flags.f_contiguous
    AO_bra = np.zeros_like(PSI_bra, order="F"); AO_ket = np.full_like(PSI_bra, np.nan, order="F")
    n = C.c_int(N); ti = C.c_double(t_init); tm = C.c_double(t_max)
    if int(os.environ.get("DYNEMOL_B200_GPUS", "1") or 1) > 1:
        _prefer_torch_nccl()
    if batched:
        tau_a = np.ascontiguousarray(np.broadcast_to(np.asarray(tau, dtype=np.float64), (2,))).copy(); save = np.zeros(2)
        lib.propagationelhl2_gpucaller_(C.byref(n), _p(S), _p(h), _p(Hp), _p(AO_bra), _p(AO_ket), _p(PSI_bra), _p(PSI_ket),
                                        C.byref(ti), C.byref(tm), _p(tau_a), _p(save))
    else:
        tau_c = C.c_double(float(tau)); sv = C.c_double(0.0)
        lib.propagationelhl_gpucaller_(C.byref(n), _p(S), _p(h), _p(Hp), _p(AO_bra), _p(AO_ket), _p(PSI_bra), _p(PSI_ket),
                                       C.byref(ti), C.byref(tm), C.byref(tau_c), C.byref(sv))
        save = np.array([sv.value])
    return dict(H_prime=Hp, AO_bra=AO_bra, AO_ket=AO_ket, PSI_bra=PSI_bra, PSI_ket=PSI_ket, save_tau=save)


def legacy_propagation(H, PSI_bra, PSI_ket, t_init, t_max, tau):
    """call Propagation_gpucaller(n, tau, save_tau, t_init, t_max, PSI_bra, PSI_ket, H) -- Taylor_gpu.cpp:295-330."""
    H = _fd(H); N = H.shape[0]
    PSI_bra = _fz(PSI_bra); PSI_ket = _fz(PSI_ket)
    n = C.c_int(N); tau_c = C.c_double(float(tau)); sv = C.c_double(0.0)
    lib.propagation_gpucaller_(C.byref(n), C.byref(tau_c), C.byref(sv), _ref(t_init, C.c_double), _ref(t_max, C.c_double),
                               _p(PSI_bra), _p(PSI_ket), _p(H))
    return PSI_bra, PSI_ket, sv.value


def legacy_ehrenfestkernel(H, A, X):
    """call EhrenfestKernel_gpu( N, H_prime, A_ad_nd, X_ij, Kernel ) -- diabatic-Ehren.f:115."""
    H = _fd(H); A = _fd(A); X = _fd(X); N = H.shape[0]
    K = np.empty((N, N), dtype=np.float64, order="F")
    lib.ehrenfestkernel_gpu_(C.byref(C.c_int(N)), _p(H), _p(A), _p(X), _p(K))
    return K


def legacy_ehrenfestkernel2(bra, ket, H, X):
    """ehrenfestkernel2_gpu_(N, bra, ket, H_prime, X_ij, Kernel) -- Taylor_gpu.cpp:801-873; bra, ket: (N, 2) complex AO packets."""
    H = _fd(H); X = _fd(X); bra = _fz(bra); ket = _fz(ket); N = H.shape[0]
    assert bra.shape == (N, 2) and ket.shape == (N, 2)
    K = np.empty((N, N), dtype=np.float64, order="F")
    lib.ehrenfestkernel2_gpu_(C.byref(C.c_int(N)), _p(bra), _p(ket), _p(H), _p(X), _p(K))
    return K


def nakedbessel(n: int, x: float) -> float:
    return lib.nakedbessel_(C.byref(C.c_int(n)), C.byref(C.c_double(x)))


def gpu_init(pid: int = 0, procs_per_dev: int = 1):
    lib.gpu_init_(C.byref(C.c_int(pid)), C.byref(C.c_int(procs_per_dev)))


def legacy_passes_last() -> int:
    """el+hole terms of the last propagation made through a legacy symbol."""
    return int(lib.dyb_legacy_passes_last())


def gpu_finalize():
    lib.gpu_finalize_()


def gpu_pin(a: np.ndarray):
    """call GPU_Pin(a, n*n*8): the size travels as a Fortran DEFAULT integer, i.e. wrapped to 32 bits, like the caller's."""
    lib.gpu_pin_(_p(a), C.byref(C.c_int(int(np.array(a.nbytes, dtype=np.int64).astype(np.int32)))))


def gpu_unpin(a: np.ndarray):
    lib.gpu_unpin_(_p(a))


def xpu_syinvert(A, uplo: str = "U"):
    """call xPU_syInvert(A, UpLo, N, info) -- GPU_Interface.cpp:861-873 (Matrix_math.f:193).  Returns (A^-1, info)."""
    A = np.array(A, dtype=np.float64, order="F", copy=True); n = C.c_int(A.shape[0]); info = C.c_int(0)
    lib.xpu_syinvert_(_p(A), C.c_char_p(uplo.encode()), C.byref(n), C.byref(info))
    return A, info.value


def xpu_dsymm(A, B, side: str = "L", uplo: str = "U", alpha: float = 1.0, beta: float = 0.0, Cin=None):
    """call xPU_dsymm(side, uplo, m, n, alpha, A, ldA, B, ldB, beta, C, ldC) -- GPU_Interface.cpp:574-627 (Matrix_math.f:119)."""
    A = _fd(A); B = _fd(B); m, n = B.shape
    out = np.zeros((m, n), dtype=np.float64, order="F") if Cin is None else np.array(Cin, dtype=np.float64, order="F", copy=True)
    lib.xpu_dsymm_(C.c_char_p(side.encode()), C.c_char_p(uplo.encode()), C.byref(C.c_int(m)), C.byref(C.c_int(n)), C.byref(C.c_double(alpha)),
                   _p(A), C.byref(C.c_int(A.shape[0])), _p(B), C.byref(C.c_int(m)), C.byref(C.c_double(beta)), _p(out), C.byref(C.c_int(m)))
    return out


def xpu_dzgemv(trans: str, A, x, alpha=1.0 + 0.0j, beta=0.0 + 0.0j, y=None):
    """call xPU_dzgemv(trans, m, n, alpha, A, ldA, x, 1, beta, y, 1) -- GPU_Interface.cpp:405-496 (Matrix_math.f:234-300)."""
    A = _fd(A); m, n = A.shape
    x = _fz(x)
    ly = n if trans.upper() in ("T", "C") else m
    out = np.zeros(ly, dtype=np.complex128) if y is None else _fz(y)
    a = np.array([alpha.real, alpha.imag]); b = np.array([complex(beta).real, complex(beta).imag])
    one = C.c_int(1)
    lib.xpu_dzgemv_(C.c_char_p(trans.encode()), C.byref(C.c_int(m)), C.byref(C.c_int(n)), _p(a), _p(A), C.byref(C.c_int(m)),
                    _p(x), C.byref(one), _p(b), _p(out), C.byref(one))
    return out
