// ============================================================================
// oracle/elhl_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17 + OpenMP, no dependencies) of the reference's CPU
// electron-hole wavepacket propagator.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
// The product path (dynemol_b200/csrc) never links or calls it.
//
// PARITY STATUS: "parity unpinned" for the Fortran part.  The reference ships
// no tests, golden vectors or known-answer files for this path (SURVEY.md
// F3, section 8c) and its CPU path (Fortran 2008 + Intel MKL + Intel MPI) cannot be
// built in this image (no Fortran compiler).  What IS pinned:
//   * orc_naked_bessel  against the reference's own nakedbessel_ compiled
//     from /root/reference/Chebyshev_gpu.cpp (oracle/_ref, see Makefile);
//   * orc_sy_invert / orc_sy_multiply against the reference's own
//     xpu_syinvert_/xpu_dsymm_ compiled from /root/reference/GPU_Interface.cpp
//     in CPU mode on top of OpenBLAS' dsytrf/dsytri/dsymm (oracle/_ref);
//   * the whole restatement against an independently written numpy
//     transcription (oracle/taylor_numpy.py) and against scipy expm.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Fortran arrays are 1-based; C(k) here is stored at C[k-1].
// ============================================================================
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef std::complex<double> cplx;

// Taylor.f:20-22 (module parameters) and constants_m.f:23
static const int    ORDER      = 25;
static const double ERROR_TOL  = 1.0e-8;
static const double NORM_ERROR = 1.0e-8;
static const double H_BAR      = 6.58264e-4;   // eV * ps

// ---------------------------------------------------------------------------
// decision trace: one record per Convergence call / steady sub-step, so that
// parity tests can compare *decisions* (tau schedule, exit index) and not only
// final vectors (SURVEY.md section 7 "hard parts").
// ---------------------------------------------------------------------------
struct orc_trace {
    int32_t n_convergence_calls;   // total Convergence() invocations
    int32_t n_substeps;            // passes through the `do while(t<t_max)` loop
    int32_t n_matvec_pairs;        // (bra_x_op, op_x_ket) pairs executed
    int32_t n_rescale;             // "rescaling tau" events (tau*0.975)
    int32_t n_first_shrink;        // tau*0.9 events in the first loop
    int32_t last_k_ref;            // k_ref after the last Convergence
    int32_t n_events;              // number of valid entries below
    int32_t ev_kind[256];          // 1 = Convergence, 2 = steady sub-step
    int32_t ev_k[256];             // Convergence: exit k (0 if failed); steady: k_ref
    int32_t ev_ok[256];            // success flag
    double  ev_tau[256];           // tau used
    double  norm_ref;
    double  final_tau;
};

static void trace_event(orc_trace* tr, int kind, int k, int ok, double tau) {
    if (!tr) return;
    if (tr->n_events < 256) {
        int e = tr->n_events;
        tr->ev_kind[e] = kind; tr->ev_k[e] = k; tr->ev_ok[e] = ok; tr->ev_tau[e] = tau;
    }
    tr->n_events++;
}

// ---------------------------------------------------------------------------
// Matrix_math.f:238-248, 291-301  (vec_bra_x_Op_alpha / vec_Op_x_ket_alpha):
//   MKL dzgemv(trans, n, n, alpha, Op, n, x, 1, (0,0), res, 1)
// real col-major matrix times complex vector, complex alpha.  'T' is a plain
// transpose (no conjugation: the matrix is real).
// ---------------------------------------------------------------------------
static void dzgemv(char trans, int n, cplx alpha, const double* A, int lda,
                   const cplx* x, cplx* y)
{
    if (trans == 'T' || trans == 't') {
        // y_j = alpha * sum_i A(i,j) x_i      (columns are contiguous -> dot products)
        #pragma omp parallel for schedule(static)
        for (int j = 0; j < n; ++j) {
            const double* col = A + (size_t)j * lda;
            double sr = 0.0, si = 0.0;
            for (int i = 0; i < n; ++i) {
                sr += col[i] * x[i].real();
                si += col[i] * x[i].imag();
            }
            y[j] = alpha * cplx(sr, si);
        }
    } else {
        // y_i = alpha * sum_j A(i,j) x_j      (axpy over columns, row-blocked per thread)
        #pragma omp parallel
        {
            int nt = 1, id = 0;
#ifdef _OPENMP
            nt = omp_get_num_threads(); id = omp_get_thread_num();
#endif
            int blk = (n + nt - 1) / nt;
            int i0 = std::min(n, id * blk), i1 = std::min(n, i0 + blk);
            const int CH = 512;                       // keep the accumulators in L1
            std::vector<double> yr(CH), yi(CH);
            for (int ib = i0; ib < i1; ib += CH) {
                int ie = std::min(i1, ib + CH), len = ie - ib;
                std::fill(yr.begin(), yr.begin() + len, 0.0);
                std::fill(yi.begin(), yi.begin() + len, 0.0);
                for (int j = 0; j < n; ++j) {
                    const double* col = A + (size_t)j * lda + ib;
                    const double xr = x[j].real(), xi = x[j].imag();
                    for (int i = 0; i < len; ++i) {
                        yr[i] += col[i] * xr;
                        yi[i] += col[i] * xi;
                    }
                }
                for (int i = 0; i < len; ++i) y[ib + i] = alpha * cplx(yr[i], yi[i]);
            }
        }
    }
}

// blas95 dotc (Taylor.f:62,103,197): sum_i conj(x_i) * y_i
static cplx dotc(int n, const cplx* x, const cplx* y) {
    double sr = 0.0, si = 0.0;
    for (int i = 0; i < n; ++i) {
        // conj(x)*y = (xr*yr + xi*yi) + i (xr*yi - xi*yr)
        sr += x[i].real() * y[i].real() + x[i].imag() * y[i].imag();
        si += x[i].real() * y[i].imag() - x[i].imag() * y[i].real();
    }
    return cplx(sr, si);
}

// Taylor.f:290-303  isConverged: false as soon as abs(a(i)-b(i)) > tol
static bool is_converged(int n, const cplx* a, const cplx* b, double tol) {
    for (int i = 0; i < n; ++i)
        if (std::abs(a[i] - b[i]) > tol) return false;
    return true;
}

extern "C" {

// ---------------------------------------------------------------------------
// Taylor.f:224-239  coefficient(tau,k_max):  c(1)=1 ; c(k) = -zi*c(k-1)*(tau/(k-1))
// ---------------------------------------------------------------------------
void orc_coefficient(double tau, int k_max, cplx* C) {
    const cplx zi(0.0, 1.0);
    C[0] = cplx(1.0, 0.0);
    for (int k = 2; k <= k_max; ++k)
        C[k - 1] = -zi * C[k - 2] * (tau / (double)(k - 1));
}

// ---------------------------------------------------------------------------
// Taylor.f:132-219  Convergence(Psi_bra, Psi_ket, C, k_ref, tau, H_prime, norm_ref, OK)
// returns OK (1/0).  k_exit (optional) = 1-based k at which the series exited.
// ---------------------------------------------------------------------------
int orc_convergence(int N, const double* H, int ldh, cplx* Psi_bra, cplx* Psi_ket,
                    cplx* C, int* k_ref, double tau, double norm_ref,
                    int* k_exit, orc_trace* tr)
{
    std::vector<cplx> bra_prev(N), ket_prev(N), bra_k(N), ket_k(N);
    std::vector<cplx> old_b(N), old_k(N), new_b(N), new_k(N);
    int ok = 0, kx = 0;
    if (k_exit) *k_exit = 0;

    orc_coefficient(tau, ORDER, C);                       // Taylor.f:163

    int k_max = ORDER;                                    // Taylor.f:165-171
    for (int k = 2; k <= ORDER; ++k)
        if (std::abs(C[k - 1]) < 1.0e-16) { k_max = k; break; }
    *k_ref = k_max;                                       // Taylor.f:173

    std::copy(Psi_bra, Psi_bra + N, bra_prev.begin());    // Taylor.f:176-180
    std::copy(Psi_ket, Psi_ket + N, ket_prev.begin());
    old_b = bra_prev; old_k = ket_prev;

    for (int k = 2; k <= k_max; ++k) {                    // Taylor.f:182
        const cplx r = C[k - 1] / C[k - 2];               // Taylor.f:185
        dzgemv('T', N, r, H, ldh, bra_prev.data(), bra_k.data());   // bra_x_op, Taylor.f:186
        dzgemv('N', N, r, H, ldh, ket_prev.data(), ket_k.data());   // op_x_ket, Taylor.f:187
        if (tr) tr->n_matvec_pairs++;
        for (int i = 0; i < N; ++i) {                     // Taylor.f:190-191
            new_b[i] = old_b[i] + bra_k[i];
            new_k[i] = old_k[i] + ket_k[i];
        }
        if (is_converged(N, new_b.data(), old_b.data(), ERROR_TOL) &&          // Taylor.f:194
            is_converged(N, new_k.data(), old_k.data(), ERROR_TOL)) {          // Taylor.f:195
            const double norm_tmp = std::abs(dotc(N, new_b.data(), new_k.data()));  // :197
            if (std::fabs(norm_tmp - norm_ref) < NORM_ERROR) {                 // Taylor.f:199
                std::copy(new_b.begin(), new_b.end(), Psi_bra);
                std::copy(new_k.begin(), new_k.end(), Psi_ket);
                ok = 1; kx = k;
                if (k_exit) *k_exit = k;
                break;
            }
        }
        old_b.swap(new_b); old_k.swap(new_k);             // Taylor.f:209-210
        bra_prev.swap(bra_k); ket_prev.swap(ket_k);
    }
    if (tr) { tr->n_convergence_calls++; tr->last_k_ref = k_max; }
    trace_event(tr, 1, kx, ok, tau);
    return ok;
}

// ---------------------------------------------------------------------------
// Taylor.f:35-127  Propagation(N, H_prime, Psi_t_bra, Psi_t_ket, t_init, t_max, tau, save_tau)
// `tau` is intent(inout) in the reference (the caller's copy is modified).
// ---------------------------------------------------------------------------
void orc_propagation(int N, const double* H, int ldh, cplx* Psi_bra, cplx* Psi_ket,
                     double t_init, double t_max, double* tau_io, double* save_tau,
                     orc_trace* tr)
{
    std::vector<cplx> C(ORDER);
    std::vector<cplx> term_b(N), term_k(N), nxt_b(N), nxt_k(N), tmp_b(N), tmp_k(N);
    double tau = *tau_io;
    int k_ref = 0;
    if (tr) std::memset(tr, 0, sizeof(*tr));

    const double norm_ref = std::abs(dotc(N, Psi_bra, Psi_ket));      // Taylor.f:62
    if (tr) tr->norm_ref = norm_ref;

    for (;;) {                                                         // Taylor.f:65-70
        int ok = orc_convergence(N, H, ldh, Psi_bra, Psi_ket, C.data(), &k_ref, tau, norm_ref, nullptr, tr);
        if (ok) break;
        tau *= 0.9;
        if (tr) tr->n_first_shrink++;
    }
    *save_tau = tau;                                                   // Taylor.f:71

    double t = t_init + tau * H_BAR;                                   // Taylor.f:73
    if (t_max - t < tau * H_BAR) {                                     // Taylor.f:75-78
        tau = (t_max - t) / H_BAR;
        orc_coefficient(tau, ORDER, C.data());
    }

    while (t < t_max) {                                                // Taylor.f:81
        std::copy(Psi_bra, Psi_bra + N, term_b.begin());               // Taylor.f:84-88
        std::copy(Psi_ket, Psi_ket + N, term_k.begin());
        tmp_b = term_b; tmp_k = term_k;
        for (int k = 2; k <= k_ref; ++k) {                             // Taylor.f:90
            const cplx r = C[k - 1] / C[k - 2];                        // Taylor.f:93
            dzgemv('T', N, r, H, ldh, term_b.data(), nxt_b.data());    // Taylor.f:94
            dzgemv('N', N, r, H, ldh, term_k.data(), nxt_k.data());    // Taylor.f:95
            if (tr) tr->n_matvec_pairs++;
            for (int i = 0; i < N; ++i) { tmp_b[i] += nxt_b[i]; tmp_k[i] += nxt_k[i]; }   // :98-99
            term_b.swap(nxt_b); term_k.swap(nxt_k);
        }
        const double norm_test = std::abs(dotc(N, tmp_b.data(), tmp_k.data()));   // Taylor.f:103
        if (tr) tr->n_substeps++;
        if (std::fabs(norm_test - norm_ref) < NORM_ERROR) {            // Taylor.f:104
            std::copy(tmp_b.begin(), tmp_b.end(), Psi_bra);
            std::copy(tmp_k.begin(), tmp_k.end(), Psi_ket);
            trace_event(tr, 2, k_ref, 1, tau);
        } else {
            trace_event(tr, 2, k_ref, 0, tau);
            int ok = 0;                                                // Taylor.f:108-113
            while (!ok) {
                tau *= 0.975;
                if (tr) tr->n_rescale++;
                ok = orc_convergence(N, H, ldh, Psi_bra, Psi_ket, C.data(), &k_ref, tau, norm_ref, nullptr, tr);
            }
        }
        t += tau * H_BAR;                                              // Taylor.f:116
        if (t_max - t < tau * H_BAR) {                                 // Taylor.f:118-121
            tau = (t_max - t) / H_BAR;
            orc_coefficient(tau, ORDER, C.data());
        }
    }
    *tau_io = tau;
    if (tr) tr->final_tau = tau;
}

// ---------------------------------------------------------------------------
// A fixed number of series terms with no decisions: the timing kernel of the
// CPU baseline (same unit of work as bench.py: one el+hole term = ket 'N' and
// bra 'T' products for both particles = four dzgemv calls, like the reference
// which runs el and hole in separate MPI ranks, ElHl_Chebyshev.f:228,253).
// bra/ket hold n_part columns of length N.
// ---------------------------------------------------------------------------
void orc_terms(int N, const double* H, int ldh, int n_part, cplx* bra, cplx* ket,
               double tau, int n_terms)
{
    std::vector<cplx> C(ORDER);
    orc_coefficient(tau, ORDER, C.data());
    std::vector<cplx> yb(N), yk(N);
    for (int t = 0; t < n_terms; ++t) {
        int k = 2 + (t % (ORDER - 1));
        const cplx r = C[k - 1] / C[k - 2];
        for (int p = 0; p < n_part; ++p) {
            dzgemv('T', N, r, H, ldh, bra + (size_t)p * N, yb.data());
            dzgemv('N', N, r, H, ldh, ket + (size_t)p * N, yk.data());
            std::copy(yb.begin(), yb.end(), bra + (size_t)p * N);
            std::copy(yk.begin(), yk.end(), ket + (size_t)p * N);
        }
    }
}

// raw products for kernel-level parity (Matrix_math.f:238-301)
void orc_dzgemv(char trans, int N, const double* alpha_reim, const double* A, int lda,
                const cplx* x, cplx* y)
{
    dzgemv(trans, N, cplx(alpha_reim[0], alpha_reim[1]), A, lda, x, y);
}

// ---------------------------------------------------------------------------
// Matrix_math.f:183-198 syInvert(A, full=.true.):  xPU_syInvert('U') followed by
// Matrix_Symmetrize(A,'U') (copy the upper triangle onto the lower one,
// Matrix_math.f:125-174).  The reference factorises with LAPACK dsytrf/dsytri
// (GPU_Interface.cpp:936-949).  Restated with a dense LDL^T-free route: Gauss-
// Jordan with partial pivoting on the full symmetric matrix, then only the
// upper triangle of the result is kept and mirrored, exactly as the reference
// discards whatever dsytri leaves below the diagonal.
// returns 0 on success, k>0 if a zero pivot is met at column k.
// ---------------------------------------------------------------------------
int orc_sy_invert(int n, double* A, int lda)
{
    std::vector<double> M((size_t)n * n), Inv((size_t)n * n, 0.0);
    // read only the upper triangle (UpLo='U'), as dsytrf does
    for (int j = 0; j < n; ++j)
        for (int i = 0; i <= j; ++i) {
            M[(size_t)i + (size_t)j * n] = A[(size_t)i + (size_t)j * lda];
            M[(size_t)j + (size_t)i * n] = A[(size_t)i + (size_t)j * lda];
        }
    for (int i = 0; i < n; ++i) Inv[(size_t)i + (size_t)i * n] = 1.0;
    // work on rows of the transposed storage: M is symmetric at the start so
    // treating columns as rows is legitimate and makes the inner loops contiguous.
    for (int k = 0; k < n; ++k) {
        int piv = k; double best = std::fabs(M[(size_t)k * n + k]);
        for (int r = k + 1; r < n; ++r) {
            double v = std::fabs(M[(size_t)r * n + k]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0) return k + 1;
        if (piv != k) {
            for (int c = 0; c < n; ++c) {
                std::swap(M[(size_t)k * n + c], M[(size_t)piv * n + c]);
                std::swap(Inv[(size_t)k * n + c], Inv[(size_t)piv * n + c]);
            }
        }
        const double d = 1.0 / M[(size_t)k * n + k];
        for (int c = 0; c < n; ++c) { M[(size_t)k * n + c] *= d; Inv[(size_t)k * n + c] *= d; }
        #pragma omp parallel for schedule(static)
        for (int r = 0; r < n; ++r) {
            if (r == k) continue;
            const double f = M[(size_t)r * n + k];
            if (f == 0.0) continue;
            double* mr = &M[(size_t)r * n]; const double* mk = &M[(size_t)k * n];
            double* ir = &Inv[(size_t)r * n]; const double* ik = &Inv[(size_t)k * n];
            for (int c = 0; c < n; ++c) { mr[c] -= f * mk[c]; ir[c] -= f * ik[c]; }
        }
    }
    // Inv (row r, col c) stored at Inv[r*n+c]; keep the upper triangle and mirror it
    for (int j = 0; j < n; ++j)
        for (int i = 0; i <= j; ++i) {
            const double v = Inv[(size_t)i * n + j];
            A[(size_t)i + (size_t)j * lda] = v;
            A[(size_t)j + (size_t)i * lda] = v;
        }
    return 0;
}

// ---------------------------------------------------------------------------
// Matrix_math.f:79-121 syMultiply(A,B,C) -> xPU_dsymm('L','U',m,n,1,A,B,0,C):
// C = A*B with A symmetric, only its upper triangle referenced.
// ---------------------------------------------------------------------------
void orc_sy_multiply(int n, const double* A, int lda, const double* B, int ldb, double* C, int ldc)
{
    #pragma omp parallel for schedule(static)
    for (int j = 0; j < n; ++j) {
        double* cj = C + (size_t)j * ldc;
        const double* bj = B + (size_t)j * ldb;
        for (int i = 0; i < n; ++i) cj[i] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double b = bj[k];
            // column k of the symmetric A: rows <= k from the stored upper part of
            // column k, rows > k from row k of the upper part (A(k,i), i>k)
            const double* ak = A + (size_t)k * lda;
            for (int i = 0; i <= k; ++i) cj[i] += ak[i] * b;
            for (int i = k + 1; i < n; ++i) cj[i] += A[(size_t)k + (size_t)i * lda] * b;
        }
    }
}

// ---------------------------------------------------------------------------
// hamiltonians.f:33-63  X_ij  and  ElHl_Chebyshev.f:296-323  Build_Huckel:
//   h(i,j) = X_ij * S(i,j),  loops j=1..N, i=1..j, mirrored.
// ---------------------------------------------------------------------------
double orc_x_ij(int i, int j, const double* IP, const double* k_WH, const double* V_shift)
{
    if (i == j) return IP[i] + V_shift[i];
    const double c1 = IP[i] - IP[j];
    const double c2 = IP[i] + IP[j];
    const double c3 = (c1 / c2) * (c1 / c2);
    const double c4 = (V_shift[i] + V_shift[j]) * 0.5;
    const double kwh = (k_WH[i] + k_WH[j]) * 0.5;
    const double k_eff = kwh + c3 + c3 * c3 * (1.0 - kwh);
    return k_eff * c2 * 0.5 + c4;
}

void orc_build_huckel(int n, const double* IP, const double* k_WH, const double* V_shift,
                      const double* S, int lds, double* h, int ldh)
{
    #pragma omp parallel for schedule(dynamic, 16)
    for (int j = 0; j < n; ++j)
        for (int i = 0; i <= j; ++i) {
            const double v = orc_x_ij(i, j, IP, k_WH, V_shift) * S[(size_t)i + (size_t)j * lds];
            h[(size_t)i + (size_t)j * ldh] = v;
            h[(size_t)j + (size_t)i * ldh] = v;
        }
}

// ---------------------------------------------------------------------------
// data_output.f:242-263 pop_Slater and data_output.f:87-147 Populations_mtx:
//   pop(f,n) = Re sum_{i: fragment(i)==f} bra(i,n)*ket(i,n)   (no conjugation here:
//   the caller already passes DUAL_bra = conj(Psi_ket), ElHl_Chebyshev.f:270)
// out is (n_frag+2) x n_part column-major: row 0 = t, rows 1..n_frag, row n_frag+1 = total.
// fragment[i] in 0..n_frag-1, or <0 for "belongs to no listed fragment".
// ---------------------------------------------------------------------------
void orc_populations(int N, int n_part, int n_frag, const int32_t* fragment,
                     const cplx* bra, const cplx* ket, double t, double* out)
{
    for (int n = 0; n < n_part; ++n) {
        double* col = out + (size_t)n * (n_frag + 2);
        col[0] = t;
        for (int f = 0; f < n_frag; ++f) {
            cplx pop(0.0, 0.0);
            for (int i = 0; i < N; ++i)
                if (fragment[i] == f) pop += bra[(size_t)n * N + i] * ket[(size_t)n * N + i];
            col[1 + f] = pop.real();
        }
        cplx tot(0.0, 0.0);
        for (int i = 0; i < N; ++i) tot += bra[(size_t)n * N + i] * ket[(size_t)n * N + i];
        col[1 + n_frag] = tot.real();
    }
}

// ---------------------------------------------------------------------------
// ElHl_Chebyshev.f:329-371 QuasiParticleEnergies:
//   erg(n) = sum_j sum_i AO_bra(i,n)*H(i,j)*AO_ket(j,n)  ; out = (re,im) per particle
// ---------------------------------------------------------------------------
void orc_quasiparticle_energies(int N, int n_part, const cplx* AO_bra, const cplx* AO_ket,
                                const double* H, int ldh, double* out_reim)
{
    for (int n = 0; n < n_part; ++n) {
        double er = 0.0, ei = 0.0;
        #pragma omp parallel for reduction(+ : er, ei) schedule(static)
        for (int j = 0; j < N; ++j) {
            cplx acc(0.0, 0.0);
            for (int i = 0; i < N; ++i)
                acc += AO_bra[(size_t)n * N + i] * H[(size_t)i + (size_t)j * ldh] * AO_ket[(size_t)n * N + j];
            er += acc.real(); ei += acc.imag();
        }
        out_reim[2 * n] = er; out_reim[2 * n + 1] = ei;
    }
}

// ---------------------------------------------------------------------------
// ElHl_Chebyshev.f:148-291  one nuclear step of ElHl_Chebyshev for n_part (<=2)
// particles, given the host-built S and h for the current geometry:
//   :174-187  t_max, tau_max, tau = first ? tau_max : min(tau_max, 1.15*save_tau)
//   :206-210  S_inv = syInvert(S) ; H' = S_inv * h
//   :228,253  Propagation for electron (col 1) and hole (col 2)
//   :266      t = t_init + delta_t*frame_step
//   :269-276  DUAL_bra = conj(Psi_ket); DUAL_ket = Psi_bra;
//             AO_bra = conj(S_inv * Psi_bra); AO_ket = Psi_ket
// S is destroyed (returns S_inv), H_prime is output.  Psi_* are the module-
// level propagated states (N x n_part).  save_tau[2] persists between calls.
// ---------------------------------------------------------------------------
void orc_elhl_step(int N, int n_part, double* S, const double* h, double* H_prime,
                   cplx* Psi_bra, cplx* Psi_ket, cplx* AO_bra, cplx* AO_ket,
                   cplx* DUAL_bra, cplx* DUAL_ket,
                   double* t_io, double delta_t, int frame_step, int it, int first_call,
                   double* save_tau, orc_trace* tr /* n_part entries or null */)
{
    const double t_init  = *t_io;
    const double t_max   = delta_t * frame_step * (it - 1);
    const double tau_max = delta_t / H_BAR;

    orc_sy_invert(N, S, N);                       // S <- S_inv   (ElHl_Chebyshev.f:206-207)
    orc_sy_multiply(N, S, N, h, N, H_prime, N);   // ElHl_Chebyshev.f:210

    for (int p = 0; p < n_part; ++p) {
        double tau = first_call ? tau_max : save_tau[p] * 1.15;     // :182 / :237
        if (tau > tau_max) tau = tau_max;                           // :184 / :238
        orc_propagation(N, H_prime, N, Psi_bra + (size_t)p * N, Psi_ket + (size_t)p * N,
                        t_init, t_max, &tau, &save_tau[p], tr ? &tr[p] : nullptr);
    }
    *t_io = t_init + delta_t * frame_step;        // :266

    for (int p = 0; p < n_part; ++p) {
        cplx* pb = Psi_bra + (size_t)p * N; cplx* pk = Psi_ket + (size_t)p * N;
        for (int i = 0; i < N; ++i) {
            DUAL_bra[(size_t)p * N + i] = std::conj(pk[i]);          // :270
            DUAL_ket[(size_t)p * N + i] = pb[i];                     // :271
            AO_ket[(size_t)p * N + i]   = pk[i];                     // :276
        }
        // :274-275  AO_bra = conj( S_inv * Psi_bra )   (op_x_ket with the real S_inv)
        dzgemv('N', N, cplx(1.0, 0.0), S, N, pb, AO_bra + (size_t)p * N);
        for (int i = 0; i < N; ++i) AO_bra[(size_t)p * N + i] = std::conj(AO_bra[(size_t)p * N + i]);
    }
}

// ---------------------------------------------------------------------------
// Chebyshev_gpu.cpp:517  nakedBessel(n,x) = 2^(n-2) * (x^2+4) / x^n
// ---------------------------------------------------------------------------
double orc_naked_bessel(int n, double x)
{
    return (double)(1 << (n - 2)) * (x * x + 4.0) / std::pow(x, (double)n);
}

// ---------------------------------------------------------------------------
// Chebyshev_gpu.cpp:636-643 coefficient: c_0 = J_0(tau); c_k = 2 J_k(tau) * zi_pow[k+1]
// with zi_pow = {i, 1, -i, -1, ...} (Chebyshev_gpu.cpp:135-142)  =>  zi_pow[k+1] = (-i)^k.
// 0-based storage as in the C++ reference.
// ---------------------------------------------------------------------------
void orc_cheb_coefficient(double tau, int k_max, cplx* coeff)
{
    static const cplx pw[4] = { cplx(0, 1), cplx(1, 0), cplx(0, -1), cplx(-1, 0) };
    coeff[0] = cplx(jn(0, tau), 0.0);
    for (int k = 1; k < k_max; ++k) coeff[k] = (2.0 * jn(k, tau)) * pw[(k + 1) & 3];
}

// ---------------------------------------------------------------------------
// Chebyshev_gpu.cpp:524-632 convergence_gpu (the un-linked true Chebyshev series,
// restated on the CPU with the reference's own arithmetic):
//   phi_0 = psi ; phi_1 = H phi_0 ; phi_k = 2 H phi_{k-1} - phi_{k-2}
//   k_max = first k in 6..24 with |c_k * nakedBessel(k,tau)| < 1e-20, else 25
//   tmp1 = c_0 phi_0 + c_1 phi_1; for k=2..k_max-1: tmp2 = tmp1 + c_k phi_k;
//   exit when max|c_k phi_k| (Idamax element, cuCabs of it) < 1e-8 for bra and
//   ket and | |<tmp2_b|tmp2_k>| - norm_ref | < 1e-8.
// NOTE: H is NOT rescaled to [-1,1] in the reference; the series
// is only meaningful for tau*rho(H) <~ 1.  findMax picks the element holding the
// largest |re| or |im| and takes its modulus (Taylor_gpu.cpp:84-89 idiom).
// ---------------------------------------------------------------------------
static double find_max_idamax(int n, const cplx* v)
{
    int best = 0; double bv = -1.0;
    const double* d = reinterpret_cast<const double*>(v);
    for (int i = 0; i < 2 * n; ++i) { double a = std::fabs(d[i]); if (a > bv) { bv = a; best = i; } }
    return std::abs(v[best / 2]);
}

int orc_cheb_convergence(int N, const double* H, int ldh, cplx* Psi_bra, cplx* Psi_ket,
                         cplx* coeff, int* k_ref, double tau, double norm_ref, int* k_exit)
{
    std::vector<std::vector<cplx>> bra(3, std::vector<cplx>(N)), ket(3, std::vector<cplx>(N));
    std::vector<cplx> t1b(N), t1k(N), t2b(N), t2k(N), db(N), dk(N), hb(N), hk(N);
    const cplx one(1.0, 0.0);
    if (k_exit) *k_exit = 0;

    std::copy(Psi_bra, Psi_bra + N, bra[0].begin());
    std::copy(Psi_ket, Psi_ket + N, ket[0].begin());
    dzgemv('T', N, one, H, ldh, bra[0].data(), bra[1].data());       // :552
    dzgemv('N', N, one, H, ldh, ket[0].data(), ket[1].data());       // :553
    orc_cheb_coefficient(tau, ORDER, coeff);                         // :556
    for (int i = 0; i < N; ++i) {                                    // :559-562
        t1b[i] = coeff[0] * bra[0][i] + coeff[1] * bra[1][i];
        t1k[i] = coeff[0] * ket[0][i] + coeff[1] * ket[1][i];
    }
    int k_max = ORDER;                                               // :565-574
    for (int k = 6; k < ORDER; ++k)
        if (std::abs(coeff[k] * orc_naked_bessel(k, tau)) < 1.0e-20) { k_max = k; break; }
    *k_ref = k_max;

    int ok = 0;
    for (int k = 2; k < k_max; ++k) {                                // :577
        std::vector<cplx>& bk = bra[k % 3]; std::vector<cplx>& b1 = bra[(k - 1) % 3]; std::vector<cplx>& b2 = bra[(k - 2) % 3];
        std::vector<cplx>& kk = ket[k % 3]; std::vector<cplx>& k1 = ket[(k - 1) % 3]; std::vector<cplx>& k2 = ket[(k - 2) % 3];
        dzgemv('T', N, one, H, ldh, b1.data(), hb.data());           // :581-582  (alpha=2, beta=-1 form)
        dzgemv('N', N, one, H, ldh, k1.data(), hk.data());
        for (int i = 0; i < N; ++i) { bk[i] = 2.0 * hb[i] - b2[i]; kk[i] = 2.0 * hk[i] - k2[i]; }
        for (int i = 0; i < N; ++i) {                                // :588-589
            t2b[i] = t1b[i] + coeff[k] * bk[i]; db[i] = t2b[i] - t1b[i];
            t2k[i] = t1k[i] + coeff[k] * kk[i]; dk[i] = t2k[i] - t1k[i];
        }
        if (find_max_idamax(N, db.data()) < ERROR_TOL &&             // :593-601
            find_max_idamax(N, dk.data()) < ERROR_TOL) {
            const double nrm = std::abs(dotc(N, t2b.data(), t2k.data()));
            if (std::fabs(nrm - norm_ref) < NORM_ERROR) {            // :609
                std::copy(t2b.begin(), t2b.end(), Psi_bra);
                std::copy(t2k.begin(), t2k.end(), Psi_ket);
                ok = 1; if (k_exit) *k_exit = k;
                break;
            }
        }
        t1b.swap(t2b); t1k.swap(t2k);                                // :621-622
    }
    return ok;
}

// ===========================================================================
// "Chebyshev mode" of the product = the reference's un-linked Chebyshev driver
// (Chebyshev_gpu.cpp:347-485 chebyshev_gpu + :524-632 convergence_gpu) made valid
// for any tau by the spectral rescaling the reference omits:
//     Ht = (H' - ebar)/de ,  spectrum of Ht inside [-1,1]
//     exp(-i tau H') = exp(-i ebar tau) * sum_k a_k(R) T_k(Ht),  R = de*tau,
//     a_0 = J_0(R), a_k = 2 (-i)^k J_k(R)            (Chebyshev_gpu.cpp:636-643 with R for tau)
// Control flow, order (25), tolerances and the tau schedule are the reference's
// (identical to Taylor.f:35-127, which Chebyshev_gpu.cpp:396-470 mirrors).
// Deliberate choices where the un-linked GPU code and the CPU oracle disagree
// (follow the CPU): the term test uses the complex modulus
// with "> tol means not converged" (Taylor.f:290-303), not Idamax + "<".
// With ebar = 0, de = 1 this is the reference's series itself.
// ===========================================================================
static void cheb_scaled_coefficient(double tau, double ebar, double de, cplx* c)
{
    static const cplx pw[4] = { cplx(1, 0), cplx(0, -1), cplx(-1, 0), cplx(0, 1) };   // (-i)^k
    const double R = de * tau;
    const cplx ph = std::exp(cplx(0.0, -ebar * tau));
    c[0] = jn(0, R) * ph;
    for (int k = 1; k < ORDER; ++k) c[k] = (2.0 * jn(k, R)) * pw[k & 3] * ph;
}

static int cheb_scaled_kmax(const cplx* c, double R)
{
    for (int k = 6; k < ORDER; ++k)                                   // Chebyshev_gpu.cpp:565-574
        if (std::abs(c[k] * orc_naked_bessel(k, R)) < 1.0e-20) return k;
    return ORDER;
}

// y = Ht x  for op 'N' / 'T'
static void apply_ht(char op, int N, const double* H, int ldh, double ebar, double de, const cplx* x, cplx* y)
{
    dzgemv(op, N, cplx(1.0, 0.0), H, ldh, x, y);
    for (int i = 0; i < N; ++i) y[i] = (y[i] - ebar * x[i]) / de;
}

void orc_cheb_scaled_coefficient(double tau, double ebar, double de, cplx* c) { cheb_scaled_coefficient(tau, ebar, de, c); }

int orc_cheb_scaled_convergence(int N, const double* H, int ldh, cplx* Psi_bra, cplx* Psi_ket, cplx* C, int* k_ref,
                                double tau, double norm_ref, double ebar, double de, int* k_exit, orc_trace* tr)
{
    std::vector<cplx> b0(Psi_bra, Psi_bra + N), k0(Psi_ket, Psi_ket + N), b1(N), k1(N), b2(N), k2(N);
    std::vector<cplx> sb(N), sk(N), nb(N), nk(N);
    int ok = 0, kx = 0;
    if (k_exit) *k_exit = 0;
    cheb_scaled_coefficient(tau, ebar, de, C);
    const int k_max = cheb_scaled_kmax(C, de * tau);
    *k_ref = k_max;
    apply_ht('T', N, H, ldh, ebar, de, b0.data(), b1.data());
    apply_ht('N', N, H, ldh, ebar, de, k0.data(), k1.data());
    if (tr) tr->n_matvec_pairs++;
    for (int i = 0; i < N; ++i) { sb[i] = C[0] * b0[i] + C[1] * b1[i]; sk[i] = C[0] * k0[i] + C[1] * k1[i]; }
    for (int k = 2; k < k_max; ++k) {
        apply_ht('T', N, H, ldh, ebar, de, b1.data(), b2.data());
        apply_ht('N', N, H, ldh, ebar, de, k1.data(), k2.data());
        if (tr) tr->n_matvec_pairs++;
        for (int i = 0; i < N; ++i) { b2[i] = 2.0 * b2[i] - b0[i]; k2[i] = 2.0 * k2[i] - k0[i]; }
        for (int i = 0; i < N; ++i) { nb[i] = sb[i] + C[k] * b2[i]; nk[i] = sk[i] + C[k] * k2[i]; }
        if (is_converged(N, nb.data(), sb.data(), ERROR_TOL) && is_converged(N, nk.data(), sk.data(), ERROR_TOL)) {
            const double nrm = std::abs(dotc(N, nb.data(), nk.data()));
            if (std::fabs(nrm - norm_ref) < NORM_ERROR) {
                std::copy(nb.begin(), nb.end(), Psi_bra); std::copy(nk.begin(), nk.end(), Psi_ket);
                ok = 1; kx = k; if (k_exit) *k_exit = k;
                break;
            }
        }
        sb.swap(nb); sk.swap(nk);
        b0.swap(b1); b1.swap(b2); k0.swap(k1); k1.swap(k2);
    }
    if (tr) { tr->n_convergence_calls++; tr->last_k_ref = k_max; }
    trace_event(tr, 1, kx, ok, tau);
    return ok;
}

void orc_cheb_scaled_propagation(int N, const double* H, int ldh, cplx* Psi_bra, cplx* Psi_ket,
                                 double t_init, double t_max, double* tau_io, double* save_tau,
                                 double ebar, double de, orc_trace* tr)
{
    std::vector<cplx> C(ORDER);
    std::vector<cplx> b0(N), k0(N), b1(N), k1(N), b2(N), k2(N), sb(N), sk(N);
    double tau = *tau_io;
    int k_ref = 0;
    if (tr) std::memset(tr, 0, sizeof(*tr));
    const double norm_ref = std::abs(dotc(N, Psi_bra, Psi_ket));
    if (tr) tr->norm_ref = norm_ref;
    for (;;) {                                                         // Chebyshev_gpu.cpp:388-393
        int ok = orc_cheb_scaled_convergence(N, H, ldh, Psi_bra, Psi_ket, C.data(), &k_ref, tau, norm_ref, ebar, de, nullptr, tr);
        if (ok) break;
        tau *= 0.9;
        if (tr) tr->n_first_shrink++;
    }
    *save_tau = tau;                                                   // :396
    double t = t_init + tau * H_BAR;
    if (t_max - t < tau * H_BAR) { tau = (t_max - t) / H_BAR; cheb_scaled_coefficient(tau, ebar, de, C.data()); }   // :399-403
    while (t < t_max) {                                                // :407
        std::copy(Psi_bra, Psi_bra + N, b0.begin()); std::copy(Psi_ket, Psi_ket + N, k0.begin());
        apply_ht('T', N, H, ldh, ebar, de, b0.data(), b1.data());      // :418-419
        apply_ht('N', N, H, ldh, ebar, de, k0.data(), k1.data());
        if (tr) tr->n_matvec_pairs++;
        for (int i = 0; i < N; ++i) { sb[i] = C[0] * b0[i] + C[1] * b1[i]; sk[i] = C[0] * k0[i] + C[1] * k1[i]; }
        for (int k = 2; k < k_ref; ++k) {                              // :421-425
            apply_ht('T', N, H, ldh, ebar, de, b1.data(), b2.data());
            apply_ht('N', N, H, ldh, ebar, de, k1.data(), k2.data());
            if (tr) tr->n_matvec_pairs++;
            for (int i = 0; i < N; ++i) { b2[i] = 2.0 * b2[i] - b0[i]; k2[i] = 2.0 * k2[i] - k0[i]; }
            for (int i = 0; i < N; ++i) { sb[i] += C[k] * b2[i]; sk[i] += C[k] * k2[i]; }
            b0.swap(b1); b1.swap(b2); k0.swap(k1); k1.swap(k2);
        }
        const double nrm = std::abs(dotc(N, sb.data(), sk.data()));    // :443-444
        if (tr) tr->n_substeps++;
        if (std::fabs(nrm - norm_ref) < NORM_ERROR) {
            std::copy(sb.begin(), sb.end(), Psi_bra); std::copy(sk.begin(), sk.end(), Psi_ket);
            trace_event(tr, 2, k_ref, 1, tau);
        } else {
            trace_event(tr, 2, k_ref, 0, tau);
            int ok = 0;
            while (!ok) {                                              // :453-461
                tau *= 0.975;
                if (tr) tr->n_rescale++;
                ok = orc_cheb_scaled_convergence(N, H, ldh, Psi_bra, Psi_ket, C.data(), &k_ref, tau, norm_ref, ebar, de, nullptr, tr);
            }
        }
        t += tau * H_BAR;                                              // :464
        if (t_max - t < tau * H_BAR) { tau = (t_max - t) / H_BAR; cheb_scaled_coefficient(tau, ebar, de, C.data()); }
    }
    *tau_io = tau;
    if (tr) tr->final_tau = tau;
}

// introspection for the tests
int    orc_order(void)      { return ORDER; }
double orc_h_bar(void)      { return H_BAR; }
int    orc_trace_size(void) { return (int)sizeof(orc_trace); }
int    orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
// torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline legs of bench.py ask for the host's cores explicitly
void   orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

} // extern "C"
