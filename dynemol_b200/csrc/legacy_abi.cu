// legacy_abi.cu -- the Fortran-callable symbols of the reference (implicit interface: lower case,
// trailing underscore, all arguments by reference) implemented on top of the native dyb_* API.
//
//   propagationelhl_gpucaller_   Taylor_gpu.cpp:219-232,634-736   (called: ElHl_Chebyshev_GPU.f:269-272)
//   propagation_gpucaller_       Taylor_gpu.cpp:295-330
//   nakedbessel_                 Chebyshev_gpu.cpp:517-519
//   gpu_init_/finalize_/pin_/unpin_   GPU_Interface.cpp:226-302   (weak: the reference's own object wins)
//   propagationelhl2_gpucaller_  new batched el+hole form
//
// Error convention of the reference: void functions; CUDA errors are printed (SAFE(), Taylor_gpu.cpp:17-18),
// LAPACK failures exit (CHECK_INFO, GPU_Interface.cpp:58).  Here every failure prints the message and
// exits: continuing after a failed propagation would silently corrupt the trajectory.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/dynemol_b200.h"

namespace {

dyb_ctx* g_ctx = nullptr;
int g_N = 0;
int g_device = 0;

void die(const char* where) {
    fprintf(stderr, "dynemol_b200: %s failed: %s\n", where, dyb_last_error());
    fflush(stderr);
    exit(EXIT_FAILURE);
}
#define LCK(call, where) do { if ((call) != DYB_OK) die(where); } while (0)

// device buffers are owned by the library, allocated on first use and sized by the first N
// (Taylor_gpu.cpp:661-673 does the same with function-local statics); a different N re-creates them.
dyb_ctx* ctx_for(int N) {
    if (g_ctx && g_N == N) return g_ctx;
    if (g_ctx) { dyb_destroy(g_ctx); g_ctx = nullptr; }
    LCK(dyb_create(&g_ctx, g_device, N, 0, N), "dyb_create");
    g_N = N;
    return g_ctx;
}

int mode_from_env() {
    // taylor (default, Taylor.f semantics) | chebyshev | taylor_refgpu | chebyshev_refgpu (include/dynemol_b200.h)
    const char* m = getenv("DYNEMOL_B200_MODE");
    if (!m) return DYB_MODE_TAYLOR;
    const bool refgpu = strstr(m, "refgpu") != nullptr || strstr(m, "REFGPU") != nullptr;
    if (m[0] == 'c' || m[0] == 'C') return refgpu ? DYB_MODE_CHEBYSHEV_REFGPU : DYB_MODE_CHEBYSHEV;
    return refgpu ? DYB_MODE_TAYLOR_REFGPU : DYB_MODE_TAYLOR;
}

void elhl(int n_part, const int* N, const double* h_S, const double* h_h, double* h_H,
          dyb_complex* h_AO_bra, dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket,
          const double* t_init, const double* t_max, double* tau, double* save_tau)
{
    dyb_ctx* c = ctx_for(*N);
    LCK(dyb_form_hprime(c, h_S, h_h, h_H), "dyb_form_hprime");                 // Taylor_gpu.cpp:676-700
    LCK(dyb_set_packets(c, n_part, h_PSI_bra, h_PSI_ket), "dyb_set_packets");  // :681-683
    const int mode = mode_from_env();
    if (mode == DYB_MODE_CHEBYSHEV)            // the operator changes every nuclear step: re-estimate its spectral interval
        LCK(dyb_estimate_spectral_bounds(c, 24, 0.05, nullptr, nullptr), "dyb_estimate_spectral_bounds");
    LCK(dyb_propagate(c, mode, *t_init, *t_max, tau, save_tau, nullptr), "dyb_propagate");  // :707
    LCK(dyb_get_packets(c, n_part, h_PSI_bra, h_PSI_ket), "dyb_get_packets");  // :711-712
    LCK(dyb_ao_bra(c, n_part, h_AO_bra), "dyb_ao_bra");                        // :718-721
}

}  // namespace

extern "C" {

void propagationelhl_gpucaller_(const int* N, const double* h_S, const double* h_h, double* h_H,
                                dyb_complex* h_AO_bra, dyb_complex* /*h_AO_ket: never touched, Taylor_gpu.cpp:634-736*/,
                                dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket,
                                const double* t_init, const double* t_max, double* tau, double* save_tau)
{
    elhl(1, N, h_S, h_h, h_H, h_AO_bra, h_PSI_bra, h_PSI_ket, t_init, t_max, tau, save_tau);
}

void propagationelhl2_gpucaller_(const int* N, const double* h_S, const double* h_h, double* h_H,
                                 dyb_complex* h_AO_bra, dyb_complex* /*h_AO_ket*/,
                                 dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket,
                                 const double* t_init, const double* t_max, double* tau, double* save_tau)
{
    elhl(2, N, h_S, h_h, h_H, h_AO_bra, h_PSI_bra, h_PSI_ket, t_init, t_max, tau, save_tau);
}

void propagation_gpucaller_(const int* n, double* tau, double* save_tau, const double* t_init, const double* t_max,
                            dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket, const double* h_H)
{
    dyb_ctx* c = ctx_for(*n);
    LCK(dyb_upload_hprime(c, h_H, *n), "dyb_upload_hprime");                   // Taylor_gpu.cpp:317
    LCK(dyb_set_packets(c, 1, h_PSI_bra, h_PSI_ket), "dyb_set_packets");
    const int mode = mode_from_env();
    if (mode == DYB_MODE_CHEBYSHEV)
        LCK(dyb_estimate_spectral_bounds(c, 24, 0.05, nullptr, nullptr), "dyb_estimate_spectral_bounds");
    LCK(dyb_propagate(c, mode, *t_init, *t_max, tau, save_tau, nullptr), "dyb_propagate");
    LCK(dyb_get_packets(c, 1, h_PSI_bra, h_PSI_ket), "dyb_get_packets");
}

// Taylor_gpu.cpp:743-797, called from diabatic-Ehren.f:115 on the "kernel" MPI rank with H' received from rank 0
void ehrenfestkernel_gpu_(const int* N, const double* h_H, const double* h_A, const double* h_X, double* h_K)
{
    dyb_ctx* c = ctx_for(*N);
    LCK(dyb_upload_hprime(c, h_H, *N), "dyb_upload_hprime");
    LCK(dyb_ehrenfest_kernel(c, h_A, h_X, h_K), "dyb_ehrenfest_kernel");
}

// Taylor_gpu.cpp:801-873 (no caller in the reference tree): the same kernel from the AO packets, the density matrix
// rho / A being formed on the device
void ehrenfestkernel2_gpu_(const int* N, const dyb_complex* h_bra, const dyb_complex* h_ket, const double* h_H, const double* h_X, double* h_K)
{
    dyb_ctx* c = ctx_for(*N);
    LCK(dyb_upload_hprime(c, h_H, *N), "dyb_upload_hprime");
    LCK(dyb_ehrenfest_kernel2(c, h_bra, h_ket, h_X, h_K), "dyb_ehrenfest_kernel2");
}

// Chebyshev_gpu.cpp:517:  2^(n-2) * (x^2 + 4) / x^n
double nakedbessel_(const int* n, const double* x)
{
    return (double)(1 << (*n - 2)) * ((*x) * (*x) + 4.0) / pow(*x, (double)*n);
}

// GPU_Interface.cpp:226-263: bind round-robin to (pid / procs_per_dev) % devCount
__attribute__((weak)) void gpu_init_(const int* pid, const int* procs_per_dev)
{
    const int n = dyb_device_count();
    if (n <= 0) { fprintf(stderr, "dynemol_b200: gpu_init_: no CUDA device (there is no CPU fallback)\n"); exit(EXIT_FAILURE); }
    const int ppd = (procs_per_dev && *procs_per_dev > 0) ? *procs_per_dev : 1;
    g_device = ((pid ? *pid : 0) / ppd) % n;
    cudaSetDevice(g_device);
    printf("Process nr. %i using GPU device nr. %i of %i (dynemol_b200)\n", pid ? *pid : 0, g_device, n);
    fflush(stdout);
}

__attribute__((weak)) void gpu_finalize_(void)
{
    if (g_ctx) { dyb_destroy(g_ctx); g_ctx = nullptr; g_N = 0; }
}

// GPU_Interface.cpp:288-302
__attribute__((weak)) void gpu_pin_(void* ptr, int* size_bytes)
{
    cudaError_t e = cudaHostRegister(ptr, (size_t)(unsigned int)*size_bytes /* N*N*8 wraps a Fortran default integer at N=16384 */, cudaHostRegisterDefault);
    if (e != cudaSuccess) { printf("ERROR(gpu_pin_): %s\n%s\n\n", cudaGetErrorName(e), cudaGetErrorString(e)); cudaGetLastError(); }
}

__attribute__((weak)) void gpu_unpin_(void* ptr)
{
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { printf("ERROR(gpu_unpin_): %s\n%s\n\n", cudaGetErrorName(e), cudaGetErrorString(e)); cudaGetLastError(); }
}

}  // extern "C"
