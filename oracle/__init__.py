"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for the CPU restatement of the reference propagator
(oracle/elhl_oracle.cpp) and, when built, for the pieces of the *real*
reference that compile in this image (oracle/_ref/, see oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.  dynemol_b200 never does.

PARITY STATUS: "parity unpinned" for the Fortran path (Taylor.f,
ElHl_Chebyshev.f): the reference holds no golden vectors and cannot be built
here.  nakedBessel and syInvert/syMultiply ARE pinned against the reference's
own C++ (oracle/_ref).  See the header of elhl_oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_REF_CHEB = os.path.join(_HERE, "_ref", "libref_chebyshev_cpu.so")
_REF_XPU = os.path.join(_HERE, "_ref", "libref_xpu_cpu.so")
_REF_TAYLOR_GPU = os.path.join(_HERE, "_ref", "libref_taylor_gpu.so")
_REF_CHEB_GPU = os.path.join(_HERE, "_ref", "libref_chebyshev_gpu.so")

ORDER = 25
H_BAR = 6.58264e-4  # eV*ps, constants_m.f:23

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class Trace(C.Structure):
    """Mirror of struct orc_trace (elhl_oracle.cpp)."""
    _fields_ = [
        ("n_convergence_calls", C.c_int32),
        ("n_substeps", C.c_int32),
        ("n_matvec_pairs", C.c_int32),
        ("n_rescale", C.c_int32),
        ("n_first_shrink", C.c_int32),
        ("last_k_ref", C.c_int32),
        ("n_events", C.c_int32),
        ("ev_kind", C.c_int32 * 256),
        ("ev_k", C.c_int32 * 256),
        ("ev_ok", C.c_int32 * 256),
        ("ev_tau", C.c_double * 256),
        ("norm_ref", C.c_double),
        ("final_tau", C.c_double),
    ]

    def events(self):
        n = min(self.n_events, 256)
        return [(self.ev_kind[i], self.ev_k[i], self.ev_ok[i], self.ev_tau[i]) for i in range(n)]

    def summary(self):
        return dict(convergence_calls=self.n_convergence_calls, substeps=self.n_substeps,
                    matvec_pairs=self.n_matvec_pairs, rescale=self.n_rescale,
                    first_shrink=self.n_first_shrink, k_ref=self.last_k_ref,
                    norm_ref=self.norm_ref, final_tau=self.final_tau)


def build(force: bool = False) -> None:
    """Compile the checker (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "elhl_oracle.cpp")
    stale = (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(src)
    need_ref = os.path.exists("/root/reference/Chebyshev_gpu.cpp") and not (
        os.path.exists(_REF_CHEB) and os.path.exists(_REF_XPU) and os.path.exists(_REF_TAYLOR_GPU) and os.path.exists(_REF_CHEB_GPU))
    if force or stale or need_ref:
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        assert _lib.orc_trace_size() == C.sizeof(Trace), "orc_trace layout mismatch"
        _lib.orc_naked_bessel.restype = C.c_double
        _lib.orc_naked_bessel.argtypes = [C.c_int, C.c_double]
        _lib.orc_x_ij.restype = C.c_double
        _lib.orc_h_bar.restype = C.c_double
    return _lib


def _cp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _fz(a) -> np.ndarray:
    """complex128, Fortran-ordered, contiguous copy."""
    return np.array(a, dtype=np.complex128, order="F", copy=True)


def _fd(a) -> np.ndarray:
    return np.array(a, dtype=np.float64, order="F", copy=True)


# --------------------------------------------------------------------------- Taylor.f
def coefficient(tau: float, k_max: int = ORDER) -> np.ndarray:
    out = np.zeros(k_max, dtype=np.complex128)
    lib().orc_coefficient(C.c_double(tau), C.c_int(k_max), _cp(out))
    return out


def convergence(H, bra, ket, tau, norm_ref):
    """Taylor.f:132-219.  Returns (ok, bra, ket, C, k_ref, k_exit)."""
    H = _fd(H); n = H.shape[0]
    bra = _fz(bra); ket = _fz(ket)
    Cc = np.zeros(ORDER, dtype=np.complex128)
    k_ref = C.c_int(0); k_exit = C.c_int(0)
    ok = lib().orc_convergence(C.c_int(n), _cp(H), C.c_int(n), _cp(bra), _cp(ket), _cp(Cc),
                               C.byref(k_ref), C.c_double(tau), C.c_double(norm_ref),
                               C.byref(k_exit), None)
    return bool(ok), bra, ket, Cc, k_ref.value, k_exit.value


def propagation(H, bra, ket, t_init, t_max, tau):
    """Taylor.f:35-127.  Returns (bra, ket, tau_out, save_tau, Trace)."""
    H = _fd(H); n = H.shape[0]
    bra = _fz(bra); ket = _fz(ket)
    tau_io = C.c_double(tau); save_tau = C.c_double(0.0)
    tr = Trace()
    lib().orc_propagation(C.c_int(n), _cp(H), C.c_int(n), _cp(bra), _cp(ket),
                          C.c_double(t_init), C.c_double(t_max), C.byref(tau_io),
                          C.byref(save_tau), C.byref(tr))
    return bra, ket, tau_io.value, save_tau.value, tr


def terms(H, bra, ket, tau, n_terms):
    """Fixed number of el+hole series terms, no decisions (CPU timing kernel)."""
    H = np.asfortranarray(H, dtype=np.float64); n = H.shape[0]
    bra = _fz(bra); ket = _fz(ket)
    n_part = 1 if bra.ndim == 1 else bra.shape[1]
    lib().orc_terms(C.c_int(n), _cp(H), C.c_int(n), C.c_int(n_part), _cp(bra), _cp(ket),
                    C.c_double(tau), C.c_int(n_terms))
    return bra, ket


def dzgemv(trans: str, H, x, alpha=1.0 + 0.0j):
    H = np.asfortranarray(H, dtype=np.float64); n = H.shape[0]
    x = _fz(x); y = np.zeros(n, dtype=np.complex128)
    a = np.array([alpha.real, alpha.imag], dtype=np.float64)
    lib().orc_dzgemv(C.c_char(trans.encode()), C.c_int(n), _cp(a), _cp(H), C.c_int(n), _cp(x), _cp(y))
    return y


# --------------------------------------------------------------------------- Matrix_math.f
def sy_invert(S) -> np.ndarray:
    A = _fd(S); n = A.shape[0]
    info = lib().orc_sy_invert(C.c_int(n), _cp(A), C.c_int(n))
    if info != 0:
        raise np.linalg.LinAlgError(f"orc_sy_invert: zero pivot at {info}")
    return A


def sy_multiply(A, B) -> np.ndarray:
    A = _fd(A); B = _fd(B); n = A.shape[0]
    out = np.zeros((n, n), dtype=np.float64, order="F")
    lib().orc_sy_multiply(C.c_int(n), _cp(A), C.c_int(n), _cp(B), C.c_int(n), _cp(out), C.c_int(n))
    return out


# --------------------------------------------------------------------------- hamiltonians.f / ElHl_Chebyshev.f
def build_huckel(IP, k_WH, V_shift, S) -> np.ndarray:
    IP = _fd(IP); k_WH = _fd(k_WH); V_shift = _fd(V_shift); S = _fd(S); n = S.shape[0]
    h = np.zeros((n, n), dtype=np.float64, order="F")
    lib().orc_build_huckel(C.c_int(n), _cp(IP), _cp(k_WH), _cp(V_shift), _cp(S), C.c_int(n), _cp(h), C.c_int(n))
    return h


def populations(fragment, bra, ket, t, n_frag):
    bra = _fz(bra); ket = _fz(ket)
    if bra.ndim == 1:
        bra = bra[:, None]; ket = ket[:, None]
        bra = np.asfortranarray(bra); ket = np.asfortranarray(ket)
    n, n_part = bra.shape
    frag = np.ascontiguousarray(fragment, dtype=np.int32)
    out = np.zeros((n_frag + 2, n_part), dtype=np.float64, order="F")
    lib().orc_populations(C.c_int(n), C.c_int(n_part), C.c_int(n_frag), _cp(frag), _cp(bra), _cp(ket),
                          C.c_double(t), _cp(out))
    return out


def quasiparticle_energies(AO_bra, AO_ket, H):
    AO_bra = _fz(AO_bra); AO_ket = _fz(AO_ket); H = np.asfortranarray(H, dtype=np.float64)
    n, n_part = AO_bra.shape
    out = np.zeros(2 * n_part, dtype=np.float64)
    lib().orc_quasiparticle_energies(C.c_int(n), C.c_int(n_part), _cp(AO_bra), _cp(AO_ket), _cp(H), C.c_int(n), _cp(out))
    return out[0::2] + 1j * out[1::2]


class ElHlState:
    """Host-side state carried between nuclear steps by ElHl_Chebyshev.f (module
    variables Psi_t_bra/ket, save_tau(2), first_call_; ElHl_Chebyshev.f:34-38)."""

    def __init__(self, Psi_bra, Psi_ket):
        self.Psi_bra = _fz(Psi_bra); self.Psi_ket = _fz(Psi_ket)
        self.n, self.n_part = self.Psi_bra.shape
        self.save_tau = np.zeros(2, dtype=np.float64)
        self.first_call = True
        self.t = 0.0
        self.it = 1  # Chebyshev_driver.f:94-106: `it` is incremented before the call


def elhl_step(state: ElHlState, S, h, delta_t, frame_step=1):
    """ElHl_Chebyshev.f:148-291 for one nuclear step.  Returns dict of outputs."""
    n, n_part = state.n, state.n_part
    S = _fd(S); h = _fd(h)
    Hp = np.zeros((n, n), dtype=np.float64, order="F")
    AO_bra = np.zeros((n, n_part), dtype=np.complex128, order="F"); AO_ket = AO_bra.copy(order="F")
    DU_bra = AO_bra.copy(order="F"); DU_ket = AO_bra.copy(order="F")
    traces = (Trace * n_part)()
    state.it += 1
    t_io = C.c_double(state.t)
    lib().orc_elhl_step(C.c_int(n), C.c_int(n_part), _cp(S), _cp(h), _cp(Hp),
                        _cp(state.Psi_bra), _cp(state.Psi_ket), _cp(AO_bra), _cp(AO_ket),
                        _cp(DU_bra), _cp(DU_ket), C.byref(t_io), C.c_double(delta_t),
                        C.c_int(frame_step), C.c_int(state.it), C.c_int(1 if state.first_call else 0),
                        _cp(state.save_tau), traces)
    state.t = t_io.value
    state.first_call = False
    return dict(H_prime=Hp, S_inv=S, AO_bra=AO_bra, AO_ket=AO_ket, DUAL_bra=DU_bra, DUAL_ket=DU_ket,
                traces=[traces[i] for i in range(n_part)], t=state.t)


# --------------------------------------------------------------------------- Chebyshev_gpu.cpp (un-linked variant)
def naked_bessel(n: int, x: float) -> float:
    return lib().orc_naked_bessel(C.c_int(n), C.c_double(x))


def cheb_coefficient(tau: float, k_max: int = ORDER) -> np.ndarray:
    out = np.zeros(k_max, dtype=np.complex128)
    lib().orc_cheb_coefficient(C.c_double(tau), C.c_int(k_max), _cp(out))
    return out


def cheb_convergence(H, bra, ket, tau, norm_ref):
    H = _fd(H); n = H.shape[0]
    bra = _fz(bra); ket = _fz(ket)
    Cc = np.zeros(ORDER, dtype=np.complex128)
    k_ref = C.c_int(0); k_exit = C.c_int(0)
    ok = lib().orc_cheb_convergence(C.c_int(n), _cp(H), C.c_int(n), _cp(bra), _cp(ket), _cp(Cc),
                                    C.byref(k_ref), C.c_double(tau), C.c_double(norm_ref), C.byref(k_exit))
    return bool(ok), bra, ket, Cc, k_ref.value, k_exit.value


def cheb_scaled_coefficient(tau: float, ebar: float, de: float) -> np.ndarray:
    out = np.zeros(ORDER, dtype=np.complex128)
    lib().orc_cheb_scaled_coefficient(C.c_double(tau), C.c_double(ebar), C.c_double(de), _cp(out))
    return out


def cheb_scaled_propagation(H, bra, ket, t_init, t_max, tau, ebar, de):
    """Chebyshev mode of the product (rescaled Chebyshev_gpu.cpp:347-485).  Returns (bra, ket, tau_out, save_tau, Trace)."""
    H = _fd(H); n = H.shape[0]
    bra = _fz(bra); ket = _fz(ket)
    tau_io = C.c_double(tau); save_tau = C.c_double(0.0)
    tr = Trace()
    lib().orc_cheb_scaled_propagation(C.c_int(n), _cp(H), C.c_int(n), _cp(bra), _cp(ket), C.c_double(t_init), C.c_double(t_max),
                                      C.byref(tau_io), C.byref(save_tau), C.c_double(ebar), C.c_double(de), C.byref(tr))
    return bra, ket, tau_io.value, save_tau.value, tr


def num_threads() -> int:
    return lib().orc_num_threads()


def use_all_host_threads() -> int:
    """Make the OpenMP loops of the oracle use every core this process may run on, whatever OMP_NUM_THREADS says
    (torchrun exports OMP_NUM_THREADS=1 to its workers).  Returns the thread count now in effect."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(n))
    return num_threads()


# --------------------------------------------------------------------------- the real reference, where it compiles
def _openblas_path():
    import scipy
    pat = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so")
    hits = sorted(glob.glob(pat))
    return hits[0] if hits else None


_ref_cheb = None
_ref_xpu = None


def ref_available() -> bool:
    return os.path.exists(_REF_CHEB) and os.path.exists(_REF_XPU) and _openblas_path() is not None


def ref_nakedbessel(n: int, x: float) -> float:
    """nakedbessel_ compiled from /root/reference/Chebyshev_gpu.cpp:517-519."""
    global _ref_cheb
    if _ref_cheb is None:
        _ref_cheb = C.CDLL(_REF_CHEB)
        _ref_cheb.nakedbessel_.restype = C.c_double
    return _ref_cheb.nakedbessel_(C.byref(C.c_int(n)), C.byref(C.c_double(x)))


def _xpu():
    global _ref_xpu
    if _ref_xpu is None:
        C.CDLL(_openblas_path(), mode=C.RTLD_GLOBAL)
        _ref_xpu = C.CDLL(_REF_XPU)
    return _ref_xpu


def ref_sy_invert_upper(S) -> np.ndarray:
    """xpu_syinvert_('U') of /root/reference/GPU_Interface.cpp:861-873,936-949 (dsytrf+dsytri).
    Only the upper triangle of the result is meaningful (as in the reference)."""
    A = _fd(S); n = C.c_int(A.shape[0]); info = C.c_int(0)
    _xpu().xpu_syinvert_(_cp(A), C.c_char_p(b"U"), C.byref(n), C.byref(info))
    if info.value != 0:
        raise np.linalg.LinAlgError(f"xpu_syinvert_ info={info.value}")
    return A


def ref_dsymm_LU(A, B) -> np.ndarray:
    """xpu_dsymm_('L','U',...) of /root/reference/GPU_Interface.cpp:574-627."""
    A = _fd(A); B = _fd(B); n = C.c_int(A.shape[0])
    out = np.zeros_like(A, order="F")
    one = C.c_double(1.0); zero = C.c_double(0.0)
    _xpu().xpu_dsymm_(C.c_char_p(b"L"), C.c_char_p(b"U"), C.byref(n), C.byref(n), C.byref(one), _cp(A), C.byref(n),
                      _cp(B), C.byref(n), C.byref(zero), _cp(out), C.byref(n))
    return out


# --------------------------------------------------------------------------- the reference's own GPU propagator (GPU box only)
_ref_tgpu = None
_ref_tgpu_elhl_n = None


def ref_gpu_available() -> bool:
    """_ref/libref_taylor_gpu.so present (compiled from /root/reference/Taylor_gpu.cpp + dzgemv_kernels.cu by
    oracle/Makefile) and a CUDA device to run it on."""
    if not os.path.exists(_REF_TAYLOR_GPU):
        return False
    try:
        return _tgpu() is not None
    except OSError:
        return False


def _tgpu():
    global _ref_tgpu
    if _ref_tgpu is None:
        lib_ = C.CDLL(_REF_TAYLOR_GPU)            # RTLD_LOCAL: its symbol names are the product's legacy names too
        lib_.ref_gpu_init_.restype = C.c_int
        if lib_.ref_gpu_init_() != 0:
            raise OSError("ref_gpu_init_ failed (no CUDA device?)")
        _ref_tgpu = lib_
    return _ref_tgpu


_ref_cgpu = None


def ref_cheb_gpu_available() -> bool:
    """_ref/libref_chebyshev_gpu.so present (Chebyshev_gpu.cpp + the reference's kernels) and a CUDA device."""
    if not os.path.exists(_REF_CHEB_GPU):
        return False
    try:
        return _cgpu() is not None
    except OSError:
        return False


def _cgpu():
    global _ref_cgpu
    if _ref_cgpu is None:
        lib_ = C.CDLL(_REF_CHEB_GPU)              # RTLD_LOCAL, own copy of the shim's globals
        lib_.ref_gpu_init_.restype = C.c_int
        if lib_.ref_gpu_init_() != 0:
            raise OSError("ref_gpu_init_ failed (no CUDA device?)")
        _ref_cgpu = lib_
    return _ref_cgpu


def ref_gpu_propagation(H, bra, ket, t_init, t_max, tau, chebyshev: bool = False):
    """propagation_gpucaller_ of /root/reference/Taylor_gpu.cpp:295-330 -- or, with chebyshev=True, of
    /root/reference/Chebyshev_gpu.cpp:308-343 -- (one particle, H' given, host buffers).  Returns (bra, ket, save_tau)."""
    Hf = _fd(H); n = C.c_int(Hf.shape[0])
    b = _fz(bra).copy(order="F"); k = _fz(ket).copy(order="F")
    tau_ = C.c_double(tau); save = C.c_double(0.0)
    lib_ = _cgpu() if chebyshev else _tgpu()
    lib_.propagation_gpucaller_(C.byref(n), C.byref(tau_), C.byref(save), C.byref(C.c_double(t_init)), C.byref(C.c_double(t_max)),
                                _cp(b), _cp(k), _cp(Hf))
    return b, k, save.value


def ref_gpu_propagationelhl(S, h, bra, ket, t_init, t_max, tau):
    """propagationelhl_gpucaller_ of /root/reference/Taylor_gpu.cpp:634-736 (one particle per call, like one MPI rank of
    ElHl_Chebyshev_GPU.f:269-272).  The reference sizes its static device buffers by the first N it sees: one N per process.
    Returns (H_prime, AO_bra, Psi_bra, Psi_ket, save_tau)."""
    global _ref_tgpu_elhl_n
    Sf = _fd(S); hf = _fd(h); N = Sf.shape[0]
    if _ref_tgpu_elhl_n not in (None, N):
        raise ValueError(f"reference propagationelhl_gpucaller_ was first called with N={_ref_tgpu_elhl_n} (static buffers)")
    _ref_tgpu_elhl_n = N
    n = C.c_int(N)
    Hp = np.zeros((N, N), dtype=np.float64, order="F")
    b = _fz(bra).copy(order="F"); k = _fz(ket).copy(order="F")
    ao_b = np.zeros(N, dtype=np.complex128); ao_k = np.zeros(N, dtype=np.complex128)
    tau_ = C.c_double(tau); save = C.c_double(0.0)
    _tgpu().propagationelhl_gpucaller_(C.byref(n), _cp(Sf), _cp(hf), _cp(Hp), _cp(ao_b), _cp(ao_k), _cp(b), _cp(k),
                                       C.byref(C.c_double(t_init)), C.byref(C.c_double(t_max)), C.byref(tau_), C.byref(save))
    return Hp, ao_b, b, k, save.value


def ref_gpu_ehrenfestkernel(H, A, X):
    """ehrenfestkernel_gpu_ of /root/reference/Taylor_gpu.cpp:743-795: K = X o A - H' A (static buffers: one N per process)."""
    Hf = _fd(H); Af = _fd(A); Xf = _fd(X); N = Hf.shape[0]
    K = np.zeros((N, N), dtype=np.float64, order="F")
    _tgpu().ehrenfestkernel_gpu_(C.byref(C.c_int(N)), _cp(Hf), _cp(Af), _cp(Xf), _cp(K))
    return K
