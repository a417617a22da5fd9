// xpu_abi.cu -- the three xPU_* dispatchers of the reference's CPU/GPU linear-algebra shim that sit ON the propagator
// path (GPU_Interface.cpp:129-158): Matrix_math.f routes a2 (syInvert, :183-198), a3 (syMultiply, :79-121) and a7
// (bra_x_op / op_x_ket, :238-301) through them.  Exported WEAK, like gpu_init_ & co.: a build that keeps the reference's
// GPU_Interface.o (it also holds the eigen-solver dispatchers xPU_dsygvd*, xPU_dgemm, xPU_dzgemm, which are off this
// path) keeps its own definitions; a build that drops it gets these.
//
//   xpu_syinvert_(A, UpLo, N, info)     GPU_Interface.cpp:861-873   A <- A^-1, A symmetric N x N host matrix, lda = N.
//        Cholesky (potrf + potri) on the device; an S that is not numerically SPD falls back to LU (getrf + getrs on the
//        identity -- the route of the reference's GPU flavour, :880-906).  BOTH triangles are valid on return (the CPU
//        flavour fills only `UpLo`, Matrix_math.f:196 mirrors it afterwards; the GPU flavour returns the full inverse).
//   xpu_dsymm_(side, uplo, M, N, alpha, A, lda, B, ldb, beta, C, ldc)   GPU_Interface.cpp:574-627   cublasDsymm, host in/out.
//   xpu_dzgemv_(trans, M, N, alpha, A, lda, X, incx, beta, Y, incy)     GPU_Interface.cpp:405-496   y = alpha op(A) x + beta y
//        with a REAL matrix and COMPLEX vectors and scalars (MKL dzgemv semantics, Matrix_math.f:221-301).
//
// These are whole-matrix host-in/host-out calls (16-24 N^2 bytes over PCIe each): they exist so that the reference's
// CPU-path callers link and run unchanged.  The hot loop does NOT go through them -- it lives behind
// propagationelhl*_gpucaller_ with H' resident in HBM.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include "../../include/dynemol_b200.h"

namespace {

cublasHandle_t     g_blas = nullptr;
cusolverDnHandle_t g_solver = nullptr;
int g_handle_dev = -1;

void xdie(const char* where, const char* what) {
    fprintf(stderr, "dynemol_b200: %s: %s (there is no CPU fallback)\n", where, what);
    fflush(stderr);
    exit(EXIT_FAILURE);
}
#define XCK(call, where) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) xdie(where, cudaGetErrorString(e_)); } while (0)

void ensure_handles(const char* where) {
    int dev = 0;
    XCK(cudaGetDevice(&dev), where);
    if (g_blas && dev == g_handle_dev) return;
    if (g_blas) { cublasDestroy(g_blas); g_blas = nullptr; }
    if (g_solver) { cusolverDnDestroy(g_solver); g_solver = nullptr; }
    if (cublasCreate(&g_blas) != CUBLAS_STATUS_SUCCESS) xdie(where, "cublasCreate failed");
    if (cusolverDnCreate(&g_solver) != CUSOLVER_STATUS_SUCCESS) xdie(where, "cusolverDnCreate failed");
    g_handle_dev = dev;
}

struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t bytes, const char* where) { XCK(cudaMalloc(&p, bytes ? bytes : 8), where); }
    ~DevBuf() { if (p) cudaFree(p); }
    double* d() const { return static_cast<double*>(p); }
};

__global__ void xpu_mirror_kernel(int n, double* A, int upper_is_valid) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx % n), j = (int)(idx / n);
    if (upper_is_valid ? (i > j) : (i < j)) A[idx] = A[(size_t)j + (size_t)i * n];
}
__global__ void xpu_identity_kernel(int n, double* A) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)n * n) A[idx] = ((int)(idx % n) == (int)(idx / n)) ? 1.0 : 0.0;
}

bool is_upper(const char* u) { return u && (*u == 'U' || *u == 'u'); }
bool is_trans(const char* t) { return t && (*t == 'T' || *t == 't' || *t == 'C' || *t == 'c'); }

}  // namespace

extern "C" {

__attribute__((weak)) void xpu_syinvert_(double* A, const char* UpLo, const int* N, int* info)
{
    const char* W = "xpu_syinvert_";
    ensure_handles(W);
    const int n = *N;
    const size_t bytes = (size_t)n * n * 8;
    const bool upper = is_upper(UpLo);
    const cublasFillMode_t uplo = upper ? CUBLAS_FILL_MODE_UPPER : CUBLAS_FILL_MODE_LOWER;
    DevBuf dA(bytes, W), dinfo(sizeof(int), W);
    XCK(cudaMemcpy(dA.p, A, bytes, cudaMemcpyHostToDevice), W);
    int lwork = 0, lwork2 = 0, h_info = 0;
    cusolverDnDpotrf_bufferSize(g_solver, uplo, n, dA.d(), n, &lwork);
    cusolverDnDpotri_bufferSize(g_solver, uplo, n, dA.d(), n, &lwork2);
    {
        DevBuf work((size_t)(lwork > lwork2 ? lwork : lwork2) * 8, W);
        if (cusolverDnDpotrf(g_solver, uplo, n, dA.d(), n, work.d(), lwork, static_cast<int*>(dinfo.p)) != CUSOLVER_STATUS_SUCCESS) xdie(W, "cusolverDnDpotrf failed");
        XCK(cudaMemcpy(&h_info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost), W);
        if (h_info == 0) {
            if (cusolverDnDpotri(g_solver, uplo, n, dA.d(), n, work.d(), lwork2, static_cast<int*>(dinfo.p)) != CUSOLVER_STATUS_SUCCESS) xdie(W, "cusolverDnDpotri failed");
            XCK(cudaMemcpy(&h_info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost), W);
            xpu_mirror_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256>>>(n, dA.d(), upper ? 1 : 0);
            XCK(cudaMemcpy(A, dA.p, bytes, cudaMemcpyDeviceToHost), W);
            if (info) *info = h_info;
            return;
        }
    }
    // not numerically positive definite: LU with partial pivoting on the symmetrised input (GPU_Interface.cpp:880-906)
    XCK(cudaMemcpy(dA.p, A, bytes, cudaMemcpyHostToDevice), W);
    xpu_mirror_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256>>>(n, dA.d(), upper ? 1 : 0);
    DevBuf dB(bytes, W), ipiv((size_t)n * sizeof(int), W);
    xpu_identity_kernel<<<(unsigned)(((size_t)n * n + 255) / 256), 256>>>(n, dB.d());
    cusolverDnDgetrf_bufferSize(g_solver, n, n, dA.d(), n, &lwork);
    DevBuf work((size_t)lwork * 8, W);
    if (cusolverDnDgetrf(g_solver, n, n, dA.d(), n, work.d(), static_cast<int*>(ipiv.p), static_cast<int*>(dinfo.p)) != CUSOLVER_STATUS_SUCCESS) xdie(W, "cusolverDnDgetrf failed");
    XCK(cudaMemcpy(&h_info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost), W);
    if (h_info == 0) {
        if (cusolverDnDgetrs(g_solver, CUBLAS_OP_N, n, n, dA.d(), n, static_cast<int*>(ipiv.p), dB.d(), n, static_cast<int*>(dinfo.p)) != CUSOLVER_STATUS_SUCCESS) xdie(W, "cusolverDnDgetrs failed");
        XCK(cudaMemcpy(A, dB.p, bytes, cudaMemcpyDeviceToHost), W);
    }
    if (info) *info = h_info;
    if (h_info != 0) { fprintf(stderr, "dynemol_b200: xpu_syinvert_: matrix is singular (getrf info = %d)\n", h_info); exit(EXIT_FAILURE); }   // CHECK_INFO, GPU_Interface.cpp:58
}

__attribute__((weak)) void xpu_dsymm_(const char* side, const char* UpLo, const int* M, const int* N,
                                      const double* alpha, double* hA, const int* LDA, double* hB, const int* LDB,
                                      const double* beta, double* hC, const int* LDC)
{
    const char* W = "xpu_dsymm_";
    ensure_handles(W);
    const int m = *M, n = *N, lda = *LDA, ldb = *LDB, ldc = *LDC;
    const bool left = side && (*side == 'L' || *side == 'l');
    const int ka = left ? m : n;                                      // A is ka x ka
    DevBuf dA((size_t)lda * ka * 8, W), dB((size_t)ldb * n * 8, W), dC((size_t)ldc * n * 8, W);
    XCK(cudaMemcpy(dA.p, hA, (size_t)lda * ka * 8, cudaMemcpyHostToDevice), W);
    XCK(cudaMemcpy(dB.p, hB, (size_t)ldb * n * 8, cudaMemcpyHostToDevice), W);
    if (*beta != 0.0) XCK(cudaMemcpy(dC.p, hC, (size_t)ldc * n * 8, cudaMemcpyHostToDevice), W);
    if (cublasDsymm(g_blas, left ? CUBLAS_SIDE_LEFT : CUBLAS_SIDE_RIGHT, is_upper(UpLo) ? CUBLAS_FILL_MODE_UPPER : CUBLAS_FILL_MODE_LOWER,
                    m, n, alpha, dA.d(), lda, dB.d(), ldb, beta, dC.d(), ldc) != CUBLAS_STATUS_SUCCESS) xdie(W, "cublasDsymm failed");
    XCK(cudaMemcpy2D(hC, (size_t)ldc * 8, dC.p, (size_t)ldc * 8, (size_t)m * 8, n, cudaMemcpyDeviceToHost), W);
}

__attribute__((weak)) void xpu_dzgemv_(const char* transA, const int* M, const int* N, const dyb_complex* alpha, double* hA, const int* LDA,
                                       dyb_complex* hX, const int* incX, const dyb_complex* beta, dyb_complex* hY, const int* incY)
{
    const char* W = "xpu_dzgemv_";
    ensure_handles(W);
    const int m = *M, n = *N, lda = *LDA, incx = *incX, incy = *incY;
    const bool tr = is_trans(transA);
    const int lx = tr ? m : n, ly = tr ? n : m;                       // lengths of x and y
    if (incx < 1 || incy < 1) xdie(W, "negative or zero increments are not supported");
    // split x into contiguous real and imaginary parts, two real GEMVs on the device, recombine with the complex scalars
    std::vector<double> xr((size_t)2 * lx), t((size_t)2 * ly);
    for (int i = 0; i < lx; ++i) { xr[i] = hX[(size_t)i * incx].re; xr[(size_t)lx + i] = hX[(size_t)i * incx].im; }
    DevBuf dA((size_t)lda * n * 8, W), dx((size_t)2 * lx * 8, W), dy((size_t)2 * ly * 8, W);
    XCK(cudaMemcpy(dA.p, hA, (size_t)lda * n * 8, cudaMemcpyHostToDevice), W);
    XCK(cudaMemcpy(dx.p, xr.data(), (size_t)2 * lx * 8, cudaMemcpyHostToDevice), W);
    const double one = 1.0, zero = 0.0;
    // [t_re t_im] = op(A) [x_re x_im]  (a 2-column GEMM: one pass over A)
    if (cublasDgemm(g_blas, tr ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_OP_N, ly, 2, lx, &one, dA.d(), lda, dx.d(), lx, &zero, dy.d(), ly) != CUBLAS_STATUS_SUCCESS)
        xdie(W, "cublasDgemm failed");
    XCK(cudaMemcpy(t.data(), dy.p, (size_t)2 * ly * 8, cudaMemcpyDeviceToHost), W);
    const bool beta0 = (beta->re == 0.0 && beta->im == 0.0);
    for (int i = 0; i < ly; ++i) {
        const double tr_ = t[i], ti = t[(size_t)ly + i];
        dyb_complex y = {alpha->re * tr_ - alpha->im * ti, alpha->re * ti + alpha->im * tr_};
        if (!beta0) {
            const dyb_complex o = hY[(size_t)i * incy];
            y.re += beta->re * o.re - beta->im * o.im; y.im += beta->re * o.im + beta->im * o.re;
        }
        hY[(size_t)i * incy] = y;
    }
}

}  // extern "C"
