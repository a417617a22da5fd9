"""oracle/taylor_numpy.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Second, independently written transcription of the reference CPU propagator in
plain numpy (complex128).  It exists to cross-validate oracle/elhl_oracle.cpp
("a second, independently written numpy transcription to
~1e-13") and is only usable at small N.  It deliberately shares no code with
the C++ restatement: matrix products go through numpy's BLAS, the inverse
through LAPACK dsytrf/dsytri as the reference does (GPU_Interface.cpp:936-949).

References (relative to /root/reference): Taylor.f:35-303, ElHl_Chebyshev.f:174-276,
Matrix_math.f:125-198, data_output.f:242-263, hamiltonians.f:33-63.
"""
from __future__ import annotations

import numpy as np

ORDER = 25            # Taylor.f:20
ERROR = 1.0e-8        # Taylor.f:21
NORM_ERROR = 1.0e-8   # Taylor.f:22
H_BAR = 6.58264e-4    # constants_m.f:23


def coefficient(tau: float, k_max: int = ORDER) -> np.ndarray:
    """Taylor.f:224-239 (index 0 here is Fortran's c(1))."""
    c = np.zeros(k_max, dtype=np.complex128)
    c[0] = 1.0
    for k in range(1, k_max):
        c[k] = -1j * c[k - 1] * (tau / k)
    return c


def _is_converged(a, b, tol):
    """Taylor.f:290-303."""
    return not np.any(np.abs(a - b) > tol)


def convergence(Hp, bra, ket, tau, norm_ref, log=None):
    """Taylor.f:132-219.  Returns (ok, bra, ket, C, k_ref, k_exit); k are 1-based."""
    C = coefficient(tau)
    small = np.nonzero(np.abs(C[1:]) < 1.0e-16)[0]
    k_max = int(small[0]) + 2 if small.size else ORDER          # Taylor.f:165-171
    term_b = bra.copy(); term_k = ket.copy()
    old_b = bra.copy(); old_k = ket.copy()
    HT = Hp.T
    for k in range(2, k_max + 1):                                # Taylor.f:182
        r = C[k - 1] / C[k - 2]
        term_b = r * (HT @ term_b)                               # bra_x_op: dzgemv('T')
        term_k = r * (Hp @ term_k)                               # op_x_ket: dzgemv('N')
        if log is not None:
            log["matvec_pairs"] = log.get("matvec_pairs", 0) + 1
        new_b = old_b + term_b
        new_k = old_k + term_k
        if _is_converged(new_b, old_b, ERROR) and _is_converged(new_k, old_k, ERROR):
            norm_tmp = abs(np.vdot(new_b, new_k))                # dotc: conj(first) . second
            if abs(norm_tmp - norm_ref) < NORM_ERROR:
                return True, new_b, new_k, C, k_max, k
        old_b, old_k = new_b, new_k
    return False, bra, ket, C, k_max, 0


def propagation(Hp, bra, ket, t_init, t_max, tau, log=None):
    """Taylor.f:35-127.  Returns (bra, ket, tau, save_tau)."""
    if log is None:
        log = {}
    log.setdefault("events", [])
    norm_ref = abs(np.vdot(bra, ket))                            # Taylor.f:62
    while True:                                                  # Taylor.f:65-70
        ok, bra, ket, C, k_ref, k_exit = convergence(Hp, bra, ket, tau, norm_ref, log)
        log["events"].append((1, k_exit, int(ok), tau))
        if ok:
            break
        tau *= 0.9
    save_tau = tau
    t = t_init + tau * H_BAR
    if t_max - t < tau * H_BAR:                                  # Taylor.f:75-78
        tau = (t_max - t) / H_BAR
        C = coefficient(tau)
    HT = Hp.T
    while t < t_max:                                             # Taylor.f:81
        term_b = bra.copy(); term_k = ket.copy()
        sum_b = bra.copy(); sum_k = ket.copy()
        for k in range(2, k_ref + 1):
            r = C[k - 1] / C[k - 2]
            term_b = r * (HT @ term_b)
            term_k = r * (Hp @ term_k)
            log["matvec_pairs"] = log.get("matvec_pairs", 0) + 1
            sum_b = sum_b + term_b
            sum_k = sum_k + term_k
        norm_test = abs(np.vdot(sum_b, sum_k))
        if abs(norm_test - norm_ref) < NORM_ERROR:               # Taylor.f:104
            bra, ket = sum_b, sum_k
            log["events"].append((2, k_ref, 1, tau))
        else:
            log["events"].append((2, k_ref, 0, tau))
            ok = False
            while not ok:                                        # Taylor.f:108-113
                tau *= 0.975
                ok, bra, ket, C, k_ref, k_exit = convergence(Hp, bra, ket, tau, norm_ref, log)
                log["events"].append((1, k_exit, int(ok), tau))
        t += tau * H_BAR
        if t_max - t < tau * H_BAR:                              # Taylor.f:118-121
            tau = (t_max - t) / H_BAR
            C = coefficient(tau)
    return bra, ket, tau, save_tau


def sy_invert_full(S):
    """Matrix_math.f:183-198 + GPU_Interface.cpp:936-949: dsytrf/dsytri('U') then mirror U->L."""
    from scipy.linalg import lapack
    A = np.array(S, dtype=np.float64, order="F", copy=True)
    ldu, ipiv, info = lapack.dsytrf(A, lower=0)
    assert info == 0
    inv, info = lapack.dsytri(ldu, ipiv, lower=0)
    assert info == 0
    U = np.triu(inv)
    return U + np.triu(inv, 1).T


def h_prime(S, h):
    """ElHl_Chebyshev.f:206-210: H' = S_inv * h  (dsymm 'L','U')."""
    Sinv = sy_invert_full(S)
    return Sinv @ h, Sinv


def x_ij_matrix(IP, k_WH, V_shift):
    """hamiltonians.f:33-63, vectorised over (i,j)."""
    IP = np.asarray(IP, float); k_WH = np.asarray(k_WH, float); V = np.asarray(V_shift, float)
    c1 = IP[:, None] - IP[None, :]
    c2 = IP[:, None] + IP[None, :]
    c3 = (c1 / c2) * (c1 / c2)
    c4 = (V[:, None] + V[None, :]) * 0.5
    kwh = (k_WH[:, None] + k_WH[None, :]) * 0.5
    k_eff = kwh + c3 + c3 * c3 * (1.0 - kwh)
    X = k_eff * c2 * 0.5 + c4
    X[np.diag_indices_from(X)] = IP + V
    return X


def pop_slater(fragment, bra, ket, n_frag):
    """data_output.f:242-263 over fragments; returns [frag pops..., total]."""
    prod = bra * ket
    out = [np.sum(prod[fragment == f]).real for f in range(n_frag)]
    out.append(np.sum(prod).real)
    return np.array(out)


# ----------------------------------------------------------------------------- Chebyshev mode (product spec)
def cheb_coefficient(tau, ebar, de):
    """Chebyshev_gpu.cpp:636-643 with R = de*tau for tau and the phase of the spectral shift."""
    from scipy.special import jv
    R = de * tau
    k = np.arange(ORDER)
    c = 2.0 * jv(k, R) * (-1j) ** k * np.exp(-1j * ebar * tau)
    c[0] *= 0.5
    return c


def _naked_bessel(n, x):
    return float(1 << (n - 2)) * (x * x + 4.0) / x ** n      # Chebyshev_gpu.cpp:517


def cheb_convergence(Hp, bra, ket, tau, norm_ref, ebar, de, log=None):
    """Rescaled Chebyshev_gpu.cpp:524-632 with the CPU oracle's term test (see elhl_oracle.cpp)."""
    C = cheb_coefficient(tau, ebar, de)
    R = de * tau
    k_max = ORDER
    for k in range(6, ORDER):
        if abs(C[k] * _naked_bessel(k, R)) < 1.0e-20:
            k_max = k
            break
    ht = lambda x: (Hp @ x - ebar * x) / de
    htT = lambda x: (Hp.T @ x - ebar * x) / de
    b0, k0 = bra, ket
    b1, k1 = htT(b0), ht(k0)
    if log is not None:
        log["matvec_pairs"] = log.get("matvec_pairs", 0) + 1
    sb = C[0] * b0 + C[1] * b1; sk = C[0] * k0 + C[1] * k1
    for k in range(2, k_max):
        b2 = 2.0 * htT(b1) - b0; k2 = 2.0 * ht(k1) - k0
        if log is not None:
            log["matvec_pairs"] += 1
        nb = sb + C[k] * b2; nk = sk + C[k] * k2
        if _is_converged(nb, sb, ERROR) and _is_converged(nk, sk, ERROR):
            if abs(abs(np.vdot(nb, nk)) - norm_ref) < NORM_ERROR:
                return True, nb, nk, C, k_max, k
        sb, sk = nb, nk
        b0, b1, k0, k1 = b1, b2, k1, k2
    return False, bra, ket, C, k_max, 0


def cheb_propagation(Hp, bra, ket, t_init, t_max, tau, ebar, de, log=None):
    """Rescaled Chebyshev_gpu.cpp:347-485 (same control flow as Taylor.f:35-127)."""
    if log is None:
        log = {}
    log.setdefault("events", [])
    norm_ref = abs(np.vdot(bra, ket))
    while True:
        ok, bra, ket, C, k_ref, k_exit = cheb_convergence(Hp, bra, ket, tau, norm_ref, ebar, de, log)
        log["events"].append((1, k_exit, int(ok), tau))
        if ok:
            break
        tau *= 0.9
    save_tau = tau
    t = t_init + tau * H_BAR
    if t_max - t < tau * H_BAR:
        tau = (t_max - t) / H_BAR
        C = cheb_coefficient(tau, ebar, de)
    ht = lambda x: (Hp @ x - ebar * x) / de
    htT = lambda x: (Hp.T @ x - ebar * x) / de
    while t < t_max:
        b0, k0 = bra, ket
        b1, k1 = htT(b0), ht(k0)
        log["matvec_pairs"] = log.get("matvec_pairs", 0) + 1
        sb = C[0] * b0 + C[1] * b1; sk = C[0] * k0 + C[1] * k1
        for k in range(2, k_ref):
            b2 = 2.0 * htT(b1) - b0; k2 = 2.0 * ht(k1) - k0
            log["matvec_pairs"] += 1
            sb = sb + C[k] * b2; sk = sk + C[k] * k2
            b0, b1, k0, k1 = b1, b2, k1, k2
        if abs(abs(np.vdot(sb, sk)) - norm_ref) < NORM_ERROR:
            bra, ket = sb, sk
            log["events"].append((2, k_ref, 1, tau))
        else:
            log["events"].append((2, k_ref, 0, tau))
            ok = False
            while not ok:
                tau *= 0.975
                ok, bra, ket, C, k_ref, k_exit = cheb_convergence(Hp, bra, ket, tau, norm_ref, ebar, de, log)
                log["events"].append((1, k_exit, int(ok), tau))
        t += tau * H_BAR
        if t_max - t < tau * H_BAR:
            tau = (t_max - t) / H_BAR
            C = cheb_coefficient(tau, ebar, de)
    return bra, ket, tau, save_tau


# ----------------------------------------------------------------------------- the reference's GPU variant
def gpu_variant_convergence(Hp, bra, ket, tau, norm_ref):
    """Taylor_gpu.cpp:511-622 (convergence_gpu), transcribed: raw powers H^k psi, c_k applied in the update, one term
    fewer than Taylor.f (k = 1..k_max-1 with k_max the first 0-based k whose |c_k| < 1e-16), term test = modulus of the
    complex element that holds the largest |re| or |im| (cublasIdamax over 2n reals) < tol.
    Returns (ok, bra, ket, C, k_ref)."""
    C = coefficient(tau)
    k_max = ORDER
    for k in range(1, ORDER):
        if abs(C[k]) < 1.0e-16:
            k_max = k
            break
    pb, pk = bra, ket
    old_b, old_k = bra, ket

    def max_elem(d):
        flat = np.abs(d.view(np.float64))
        return abs(d[int(np.argmax(flat)) // 2])

    for k in range(1, k_max):
        pb = Hp.T @ pb
        pk = Hp @ pk
        new_b = old_b + C[k] * pb
        new_k = old_k + C[k] * pk
        if max_elem(new_b - old_b) < ERROR and max_elem(new_k - old_k) < ERROR:
            if abs(abs(np.vdot(new_b, new_k)) - norm_ref) < NORM_ERROR:
                return True, new_b, new_k, C, k_max
        old_b, old_k = new_b, new_k
    return False, bra, ket, C, k_max


def gpu_variant_propagation(Hp, bra, ket, t_init, t_max, tau):
    """Taylor_gpu.cpp:334-480 (chebyshev_gpu of the Taylor file).  Returns (bra, ket, tau, save_tau, n_rescale)."""
    norm_ref = abs(np.vdot(bra, ket))
    while True:
        ok, bra, ket, C, k_ref = gpu_variant_convergence(Hp, bra, ket, tau, norm_ref)
        if ok:
            break
        tau *= 0.9
    save_tau = tau
    t = t_init + tau * H_BAR
    if t_max - t < tau * H_BAR:
        tau = (t_max - t) / H_BAR
        C = coefficient(tau)
    n_rescale = 0
    while t < t_max:
        pb, pk = bra, ket
        sb = C[0] * pb; sk = C[0] * pk
        for k in range(1, k_ref):
            pb = Hp.T @ pb
            pk = Hp @ pk
            sb = sb + C[k] * pb; sk = sk + C[k] * pk
        if abs(abs(np.vdot(sb, sk)) - norm_ref) < NORM_ERROR:
            bra, ket = sb, sk
        else:
            ok = False
            while not ok:
                tau *= 0.975
                n_rescale += 1
                ok, bra, ket, C, k_ref = gpu_variant_convergence(Hp, bra, ket, tau, norm_ref)
        t += tau * H_BAR
        if t_max - t < tau * H_BAR:
            tau = (t_max - t) / H_BAR
            C = coefficient(tau)
    return bra, ket, tau, save_tau, n_rescale


def _max_elem(d):
    """findMax of Taylor_gpu.cpp:84-89 / Chebyshev_gpu.cpp:92-97: cublasIdamax over the 2n reals, then the modulus of the
    complex element that owns the winning component."""
    flat = np.abs(np.ascontiguousarray(d).view(np.float64))
    return abs(d[int(np.argmax(flat)) // 2])


def gpu_variant_cheb_coefficient(tau):
    """Chebyshev_gpu.cpp:636-643: c_0 = J_0(tau), c_k = 2 J_k(tau) (-i)^k (zi_pow[k+1]); NO spectral rescaling."""
    from scipy.special import jv
    k = np.arange(ORDER)
    c = 2.0 * jv(k, tau) * (-1j) ** k
    c[0] = jv(0, tau)
    return c


def gpu_variant_cheb_convergence(Hp, bra, ket, tau, norm_ref):
    """Chebyshev_gpu.cpp:524-632 (convergence_gpu), transcribed: phi_1 = H phi_0, phi_k = 2 H phi_{k-1} - phi_{k-2};
    k_max = first k in 6..24 with |c_k nakedBessel(k,tau)| < 1e-20 else 25; tests (Idamax style, strict <) from k = 2.
    Returns (ok, bra, ket, C, k_ref)."""
    C = gpu_variant_cheb_coefficient(tau)
    k_max = ORDER
    for k in range(6, ORDER):
        if abs(C[k] * _naked_bessel(k, tau)) < 1.0e-20:
            k_max = k
            break
    b0, k0 = bra, ket
    b1, k1 = Hp.T @ b0, Hp @ k0
    sb = C[0] * b0 + C[1] * b1; sk = C[0] * k0 + C[1] * k1
    for k in range(2, k_max):
        b2 = 2.0 * (Hp.T @ b1) - b0; k2 = 2.0 * (Hp @ k1) - k0
        nb = sb + C[k] * b2; nk = sk + C[k] * k2
        if _max_elem(nb - sb) < ERROR and _max_elem(nk - sk) < ERROR:
            if abs(abs(np.vdot(nb, nk)) - norm_ref) < NORM_ERROR:
                return True, nb, nk, C, k_max
        sb, sk = nb, nk
        b0, b1, k0, k1 = b1, b2, k1, k2
    return False, bra, ket, C, k_max


def gpu_variant_cheb_propagation(Hp, bra, ket, t_init, t_max, tau):
    """Chebyshev_gpu.cpp:347-485 (chebyshev_gpu).  Returns (bra, ket, tau, save_tau, n_rescale)."""
    norm_ref = abs(np.vdot(bra, ket))
    while True:
        ok, bra, ket, C, k_ref = gpu_variant_cheb_convergence(Hp, bra, ket, tau, norm_ref)
        if ok:
            break
        tau *= 0.9
    save_tau = tau
    t = t_init + tau * H_BAR
    if t_max - t < tau * H_BAR:
        tau = (t_max - t) / H_BAR
        C = gpu_variant_cheb_coefficient(tau)
    n_rescale = 0
    while t < t_max:
        b0, k0 = bra, ket
        b1, k1 = Hp.T @ b0, Hp @ k0
        sb = C[0] * b0 + C[1] * b1; sk = C[0] * k0 + C[1] * k1
        for k in range(2, k_ref):
            b2 = 2.0 * (Hp.T @ b1) - b0; k2 = 2.0 * (Hp @ k1) - k0
            sb = sb + C[k] * b2; sk = sk + C[k] * k2
            b0, b1, k0, k1 = b1, b2, k1, k2
        if abs(abs(np.vdot(sb, sk)) - norm_ref) < NORM_ERROR:
            bra, ket = sb, sk
        else:
            ok = False
            while not ok:
                tau *= 0.975
                n_rescale += 1
                ok, bra, ket, C, k_ref = gpu_variant_cheb_convergence(Hp, bra, ket, tau, norm_ref)
        t += tau * H_BAR
        if t_max - t < tau * H_BAR:
            tau = (t_max - t) / H_BAR
            C = gpu_variant_cheb_coefficient(tau)
    return bra, ket, tau, save_tau, n_rescale


# ----------------------------------------------------------------------------- single-expansion Chebyshev (product spec)
def cheb_full_coefficients(tau, ebar, de):
    """DYB_MODE_CHEBYSHEV_FULL: the coefficients of Chebyshev_gpu.cpp:636-643 on the rescaled operator, with the order taken
    from their decay instead of the reference's cap of 25: K = first k > R (R = de*tau) with 2|J_k(R)| < 1e-15."""
    from scipy.special import jv
    R = de * tau
    ph = np.exp(-1j * ebar * tau)
    C = [jv(0, R) * ph]
    k = 1
    while True:
        j = jv(k, R)
        if k > R and k >= 2 and 2.0 * abs(j) < 1.0e-15:
            break
        C.append(2.0 * j * (-1j) ** k * ph)
        k += 1
    return np.array(C)


def cheb_full_propagation(Hp, bra, ket, t_init, t_max, ebar, de):
    """One expansion over the whole interval, one norm test at the end.  Returns (bra, ket, n_terms, norm_ok)."""
    tau = (t_max - t_init) / H_BAR
    C = cheb_full_coefficients(tau, ebar, de)
    norm_ref = abs(np.vdot(bra, ket))
    ht = lambda x: (Hp @ x - ebar * x) / de
    htT = lambda x: (Hp.T @ x - ebar * x) / de
    b0, k0 = bra, ket
    b1, k1 = htT(b0), ht(k0)
    sb = C[0] * b0 + C[1] * b1; sk = C[0] * k0 + C[1] * k1
    for k in range(2, len(C)):
        b2 = 2.0 * htT(b1) - b0; k2 = 2.0 * ht(k1) - k0
        sb = sb + C[k] * b2; sk = sk + C[k] * k2
        b0, b1, k0, k1 = b1, b2, k1, k2
    ok = abs(abs(np.vdot(sb, sk)) - norm_ref) < NORM_ERROR
    return sb, sk, len(C) - 1, ok
