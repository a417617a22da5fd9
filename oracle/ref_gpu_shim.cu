// oracle/ref_gpu_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Lets the reference's OWN GPU propagator (Taylor_gpu.cpp + dzgemv_kernels.cu, compiled in place from /root/reference
// by oracle/Makefile into _ref/libref_taylor_gpu.so) run on the B200 box, so that the product can be checked against
// the reference itself and timed next to it.  The two reference files need four things from GPU_Interface.cpp, which
// cannot be built here (it needs MAGMA):
//   * the globals myHandle / cublas_default / stream[] / nStreams     (GPU_Interface.cpp:206-214, set up in :244-256)
//   * gpu_dgeInvert(dA, n, lddA, stream): in-place inverse of a general matrix (GPU_Interface.cpp:910-929, MAGMA
//     dgetrf_gpu + dgetri_gpu) -- stated here with cuSOLVER getrf + getrs on the identity (same LU route).
// This file is written from those interfaces; it contains no reference code.
#include <cuda_runtime_api.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <stdio.h>
#include <stdlib.h>

cublasHandle_t myHandle;
cudaStream_t   cublas_default;
extern const int nStreams = 3;
cudaStream_t   stream[3];

static cusolverDnHandle_t g_solver = nullptr;
static bool g_ready = false;

#define RCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "ref_gpu_shim %s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); abort(); } } while (0)

// what GPU_Init does for the propagator (GPU_Interface.cpp:244-256); device 0
extern "C" int ref_gpu_init_(void) {
    if (g_ready) return 0;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) return 3;      // no device: the caller skips
    if (cudaSetDevice(0) != cudaSuccess) return 3;
    if (cublasCreate(&myHandle) != CUBLAS_STATUS_SUCCESS) return 1;
    cublasSetAtomicsMode(myHandle, CUBLAS_ATOMICS_ALLOWED);
    cublasGetStream(myHandle, &cublas_default);
    for (int i = 0; i < nStreams; ++i) RCK(cudaStreamCreate(&stream[i]));
    if (cusolverDnCreate(&g_solver) != CUSOLVER_STATUS_SUCCESS) return 2;
    g_ready = true;
    return 0;
}

__global__ static void ref_shim_identity(double* A, int n, int ld) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)n * ld) { const int i = idx % ld, j = idx / ld; A[idx] = (i == j) ? 1.0 : 0.0; }
}

void gpu_dgeInvert(double* dA, const int n, const int lddA, cudaStream_t s) {
    // synchronous, like the MAGMA calls it stands in for
    RCK(cudaStreamSynchronize(s));
    cusolverDnSetStream(g_solver, s);
    int lwork = 0;
    cusolverDnDgetrf_bufferSize(g_solver, n, n, dA, lddA, &lwork);
    double *work = nullptr, *B = nullptr; int *ipiv = nullptr, *info = nullptr;
    RCK(cudaMalloc(&work, sizeof(double) * (size_t)lwork));
    RCK(cudaMalloc(&B, sizeof(double) * (size_t)lddA * n));
    RCK(cudaMalloc(&ipiv, sizeof(int) * n));
    RCK(cudaMalloc(&info, sizeof(int)));
    const size_t total = (size_t)n * lddA;
    ref_shim_identity<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(B, n, lddA);
    cusolverDnDgetrf(g_solver, n, n, dA, lddA, work, ipiv, info);
    cusolverDnDgetrs(g_solver, CUBLAS_OP_N, n, n, dA, lddA, ipiv, B, lddA, info);
    int h_info = 0;
    RCK(cudaMemcpyAsync(&h_info, info, sizeof(int), cudaMemcpyDeviceToHost, s));
    RCK(cudaMemcpyAsync(dA, B, sizeof(double) * total, cudaMemcpyDeviceToDevice, s));
    RCK(cudaStreamSynchronize(s));
    if (h_info != 0) fprintf(stderr, "ref_gpu_shim: getrf/getrs info = %d\n", h_info);
    cudaFree(work); cudaFree(B); cudaFree(ipiv); cudaFree(info);
}
