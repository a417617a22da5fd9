/* ==========================================================================
 * oracle/ref_blas_shim.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Link-time shim that lets the reference's own GPU_Interface.cpp (compiled in
 * its CPU mode, -DFORTRAN, no -DUSE_GPU, straight from /root/reference) resolve
 * the Fortran BLAS/LAPACK symbols it calls (GPU_Interface.cpp:97-121).  The
 * reference links Intel MKL, which is not in this image; the LP64 OpenBLAS that
 * ships inside scipy exports the same routines under a `scipy_` prefix, so the
 * standard ones are forwarded to it with dlsym().  The loader (oracle/__init__.py)
 * dlopens scipy's OpenBLAS with RTLD_GLOBAL before this library.
 *
 * dzgemv_/dzgemm_ are MKL-only extensions (real matrix x complex vector/matrix,
 * Matrix_math.f:221,271); there is nothing to forward to, so they are written
 * out here from their documented semantics.  They are only needed to satisfy
 * the linker for xpu_dzgemv_/xpu_dzgemm_; the oracle itself has its own dzgemv.
 * ========================================================================== */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

typedef struct { double re, im; } zc;

static void* must(const char* name)
{
    void* p = dlsym(RTLD_DEFAULT, name);
    if (!p) { fprintf(stderr, "ref_blas_shim: symbol %s not found (load scipy openblas RTLD_GLOBAL first)\n", name); abort(); }
    return p;
}

void dsymm_(const char* side, const char* uplo, const int* m, const int* n, const double* alpha,
            double* A, const int* lda, double* B, const int* ldb, const double* beta, double* C, const int* ldc)
{
    typedef void (*fn)(const char*, const char*, const int*, const int*, const double*, double*, const int*,
                       double*, const int*, const double*, double*, const int*);
    static fn f; if (!f) f = (fn)must("scipy_dsymm_");
    f(side, uplo, m, n, alpha, A, lda, B, ldb, beta, C, ldc);
}

void dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
            double* A, const int* lda, double* B, const int* ldb, const double* beta, double* C, const int* ldc)
{
    typedef void (*fn)(const char*, const char*, const int*, const int*, const int*, const double*, double*,
                       const int*, double*, const int*, const double*, double*, const int*);
    static fn f; if (!f) f = (fn)must("scipy_dgemm_");
    f(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

void dsytrf_(const char* uplo, const int* n, double* A, const int* lda, int* ipiv, double* work, const int* lwork, int* info)
{
    typedef void (*fn)(const char*, const int*, double*, const int*, int*, double*, const int*, int*);
    static fn f; if (!f) f = (fn)must("scipy_dsytrf_");
    f(uplo, n, A, lda, ipiv, work, lwork, info);
}

void dsytri_(const char* uplo, const int* n, double* A, const int* lda, int* ipiv, double* work, int* info)
{
    typedef void (*fn)(const char*, const int*, double*, const int*, int*, double*, int*);
    static fn f; if (!f) f = (fn)must("scipy_dsytri_");
    f(uplo, n, A, lda, ipiv, work, info);
}

void dsygvd_(const int* itype, const char* jobz, const char* uplo, const int* n, double* A, const int* lda,
             double* B, const int* ldb, double* W, double* work, const int* lwork, int* iwork, const int* liwork, int* info)
{
    typedef void (*fn)(const int*, const char*, const char*, const int*, double*, const int*, double*, const int*,
                       double*, double*, const int*, int*, const int*, int*);
    static fn f; if (!f) f = (fn)must("scipy_dsygvd_");
    f(itype, jobz, uplo, n, A, lda, B, ldb, W, work, lwork, iwork, liwork, info);
}

/* y = alpha*op(A)*x + beta*y, A real m x n col-major, x,y complex (MKL dzgemv) */
void dzgemv_(const char* trans, const int* M, const int* N, const zc* alpha, double* A, const int* lda,
             zc* x, const int* incx, const zc* beta, zc* y, const int* incy)
{
    const int m = *M, n = *N, t = (*trans == 'T' || *trans == 't' || *trans == 'C' || *trans == 'c');
    const int leny = t ? n : m, lenx = t ? m : n;
    for (int i = 0; i < leny; ++i) {
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < lenx; ++k) {
            const double a = t ? A[(size_t)k + (size_t)i * *lda] : A[(size_t)i + (size_t)k * *lda];
            sr += a * x[(size_t)k * *incx].re; si += a * x[(size_t)k * *incx].im;
        }
        zc* yy = &y[(size_t)i * *incy];
        const double br = beta->re * yy->re - beta->im * yy->im, bi = beta->re * yy->im + beta->im * yy->re;
        yy->re = alpha->re * sr - alpha->im * si + br;
        yy->im = alpha->re * si + alpha->im * sr + bi;
    }
}

/* C = alpha*op(A)*B + beta*C, A real, B,C complex (MKL dzgemm); transB must be 'N' */
void dzgemm_(const char* ta, const char* tb, const int* M, const int* N, const int* K, const zc* alpha,
             double* A, const int* lda, const zc* B, const int* ldb, const zc* beta, zc* C, const int* ldc)
{
    (void)tb;
    const int t = (*ta == 'T' || *ta == 't' || *ta == 'C' || *ta == 'c');
    for (int j = 0; j < *N; ++j)
        for (int i = 0; i < *M; ++i) {
            double sr = 0.0, si = 0.0;
            for (int k = 0; k < *K; ++k) {
                const double a = t ? A[(size_t)k + (size_t)i * *lda] : A[(size_t)i + (size_t)k * *lda];
                sr += a * B[(size_t)k + (size_t)j * *ldb].re; si += a * B[(size_t)k + (size_t)j * *ldb].im;
            }
            zc* cc = &C[(size_t)i + (size_t)j * *ldc];
            const double br = beta->re * cc->re - beta->im * cc->im, bi = beta->re * cc->im + beta->im * cc->re;
            cc->re = alpha->re * sr - alpha->im * si + br;
            cc->im = alpha->re * si + alpha->im * sr + bi;
        }
}
