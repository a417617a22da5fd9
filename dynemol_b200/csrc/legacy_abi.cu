// legacy_abi.cu -- the Fortran-callable symbols of the reference (implicit interface: lower case,
// trailing underscore, all arguments by reference) implemented on top of the native dyb_* API.
//
//   propagationelhl_gpucaller_   Taylor_gpu.cpp:219-232,634-736   (called: ElHl_Chebyshev_GPU.f:269-272)
//   propagation_gpucaller_       Taylor_gpu.cpp:295-330
//   ehrenfestkernel_gpu_         Taylor_gpu.cpp:743-797           (called: diabatic-Ehren.f:115)
//   ehrenfestkernel2_gpu_        Taylor_gpu.cpp:801-873
//   nakedbessel_                 Chebyshev_gpu.cpp:517-519
//   gpu_init_/finalize_/pin_/unpin_   GPU_Interface.cpp:226-302   (weak: the reference's own object wins)
//   propagationelhl2_gpucaller_  new batched el+hole form
// (the weak xpu_* dispatchers of GPU_Interface.cpp:129-158 live in xpu_abi.cu)
//
// Environment (read at every call):
//   DYNEMOL_B200_MODE     taylor (default, Taylor.f semantics) | chebyshev (one expansion per step, order from the Bessel
//                         decay) | chebyshev25 (the reference's order-25 series, rescaled) | taylor_refgpu | chebyshev_refgpu
//   DYNEMOL_B200_GPUS     P > 1: row-shard H' over P GPUs of this box inside the call (single-process team, team.cu)
//   DYNEMOL_B200_AUTOPIN  1: page-lock the caller's S, h, H' buffers when first seen (what GPU_Pin does in
//                         ElHl_Chebyshev_GPU.f:109-111) -- only for callers whose buffers live as long as the run
//
// Error convention of the reference: void functions; CUDA errors are printed (SAFE(), Taylor_gpu.cpp:17-18),
// LAPACK failures exit (CHECK_INFO, GPU_Interface.cpp:58).  Here every failure prints the message and
// exits: continuing after a failed propagation would silently corrupt the trajectory.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/dynemol_b200.h"

namespace {

dyb_ctx*  g_ctx = nullptr;
dyb_team* g_team = nullptr;
int g_N = 0, g_team_N = 0, g_team_P = 0;
int g_device = 0;
bool g_init_by_us = false;        // our weak gpu_init_ ran (else the reference's GPU_Init bound the device, GPU_Interface.cpp:244-245)
std::vector<void*> g_autopinned;

void die(const char* where, const char* msg) {
    fprintf(stderr, "dynemol_b200: %s failed: %s\n", where, msg);
    fflush(stderr);
    exit(EXIT_FAILURE);
}
#define LCK(call, where) do { if ((call) != DYB_OK) die(where, dyb_last_error()); } while (0)
#define TCK(call, where) do { if ((call) != DYB_OK) die(where, dyb_team_last_error()); } while (0)

// The legacy calls run on whatever device the host bound this process/rank to and leave the caller's current device
// untouched: with the reference's GPU_Interface.o linked, its strong gpu_init_ did cudaSetDevice(myGPU) and later xPU_*
// calls expect that binding to stand.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
int bound_device() {
    if (g_init_by_us) return g_device;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev;
}

// device buffers are owned by the library, allocated on first use and sized by the first N
// (Taylor_gpu.cpp:661-673 does the same with function-local statics); a different N re-creates them.
dyb_ctx* ctx_for(int N) {
    const int dev = bound_device();
    if (g_ctx && g_N == N && g_device == dev) return g_ctx;
    if (g_ctx) { dyb_destroy(g_ctx); g_ctx = nullptr; }
    if (g_team) { dyb_team_destroy(g_team); g_team = nullptr; g_team_N = g_team_P = 0; }
    g_device = dev;
    LCK(dyb_create(&g_ctx, dev, N, 0, N), "dyb_create");
    g_N = N;
    return g_ctx;
}

int team_size_from_env() {
    const char* e = getenv("DYNEMOL_B200_GPUS");
    const int p = e ? atoi(e) : 1;
    return p > 1 ? p : 1;
}
dyb_team* team_for(int N, int P) {
    if (g_team && g_team_N == N && g_team_P == P) return g_team;
    if (g_team) { dyb_team_destroy(g_team); g_team = nullptr; }
    if (g_ctx) { dyb_destroy(g_ctx); g_ctx = nullptr; g_N = 0; }       // free the single-GPU buffers of an earlier mode
    // devices bound_device() .. +P-1 (mod count): a rank bound by GPU_Init starts its team at its own device
    const int n = dyb_device_count(), d0 = bound_device();
    std::vector<int> devs(P);
    for (int r = 0; r < P; ++r) devs[r] = n > 0 ? (d0 + r) % n : r;
    TCK(dyb_team_create(&g_team, P, devs.data(), N), "dyb_team_create");
    g_team_N = N; g_team_P = P;
    return g_team;
}

struct Mode { int mode; bool bounds; };
Mode mode_from_env() {
    const char* m = getenv("DYNEMOL_B200_MODE");
    if (!m) return {DYB_MODE_TAYLOR, false};
    const bool refgpu = strstr(m, "refgpu") != nullptr || strstr(m, "REFGPU") != nullptr;
    if (m[0] == 'c' || m[0] == 'C') {
        if (refgpu) return {DYB_MODE_CHEBYSHEV_REFGPU, false};
        if (strstr(m, "25")) return {DYB_MODE_CHEBYSHEV, true};
        return {DYB_MODE_CHEBYSHEV_FULL, true};
    }
    return {refgpu ? DYB_MODE_TAYLOR_REFGPU : DYB_MODE_TAYLOR, false};
}

// opt-in: page-lock a caller buffer the first time it is seen (skipped when it already is pinned / registered)
void autopin(const void* p, size_t bytes) {
    static const bool on = getenv("DYNEMOL_B200_AUTOPIN") && getenv("DYNEMOL_B200_AUTOPIN")[0] == '1';
    if (!on || !p) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type != cudaMemoryTypeUnregistered) return;
    cudaGetLastError();
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) g_autopinned.push_back(const_cast<void*>(p));
    else cudaGetLastError();
}

constexpr int LANCZOS_STEPS = 24;
constexpr double LANCZOS_MARGIN = 0.05;

void elhl(int n_part, const int* N, const double* h_S, const double* h_h, double* h_H,
          dyb_complex* h_AO_bra, dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket,
          const double* t_init, const double* t_max, double* tau, double* save_tau)
{
    DeviceGuard guard;
    const size_t nn = (size_t)*N * (size_t)*N * 8;
    autopin(h_S, nn); autopin(h_h, nn); autopin(h_H, nn);
    const Mode md = mode_from_env();
    const int P = team_size_from_env();
    if (P > 1) {                           // single-process multi-GPU: H' row-sharded over P GPUs of this box (team.cu)
        dyb_team* t = team_for(*N, P);
        TCK(dyb_team_form_hprime(t, h_S, h_h, h_H), "dyb_team_form_hprime");
        TCK(dyb_team_set_packets(t, n_part, h_PSI_bra, h_PSI_ket), "dyb_team_set_packets");
        if (md.bounds) TCK(dyb_team_estimate_spectral_bounds(t, LANCZOS_STEPS, LANCZOS_MARGIN, nullptr, nullptr), "dyb_team_estimate_spectral_bounds");
        TCK(dyb_team_propagate(t, md.mode, *t_init, *t_max, tau, save_tau, nullptr), "dyb_team_propagate");
        TCK(dyb_team_get_packets(t, n_part, h_PSI_bra, h_PSI_ket), "dyb_team_get_packets");
        TCK(dyb_team_ao_bra(t, n_part, h_AO_bra), "dyb_team_ao_bra");
        TCK(dyb_team_wait_outputs(t), "dyb_team_wait_outputs");
        return;
    }
    dyb_ctx* c = ctx_for(*N);
    // S up, potrf queued, h up behind it, solve, H' on its way down while the series runs (Taylor_gpu.cpp:676-700 does the
    // same dance with three streams and events)
    LCK(dyb_form_hprime_async(c, h_S, h_h, h_H), "dyb_form_hprime");
    LCK(dyb_set_packets(c, n_part, h_PSI_bra, h_PSI_ket), "dyb_set_packets");  // :681-683
    if (md.bounds)                             // the operator changes every nuclear step: re-estimate its spectral interval
        LCK(dyb_estimate_spectral_bounds(c, LANCZOS_STEPS, LANCZOS_MARGIN, nullptr, nullptr), "dyb_estimate_spectral_bounds");
    LCK(dyb_propagate(c, md.mode, *t_init, *t_max, tau, save_tau, nullptr), "dyb_propagate");  // :707
    LCK(dyb_get_packets(c, n_part, h_PSI_bra, h_PSI_ket), "dyb_get_packets");  // :711-712
    LCK(dyb_ao_bra(c, n_part, h_AO_bra), "dyb_ao_bra");                        // :718-721
    LCK(dyb_wait_outputs(c), "download of H'");                                // :724-725: returns fully synchronised
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
    DeviceGuard guard;
    const Mode md = mode_from_env();
    const int P = team_size_from_env();
    if (P > 1) {
        dyb_team* t = team_for(*n, P);
        TCK(dyb_team_upload_hprime(t, h_H, *n), "dyb_team_upload_hprime");
        TCK(dyb_team_set_packets(t, 1, h_PSI_bra, h_PSI_ket), "dyb_team_set_packets");
        if (md.bounds) TCK(dyb_team_estimate_spectral_bounds(t, LANCZOS_STEPS, LANCZOS_MARGIN, nullptr, nullptr), "dyb_team_estimate_spectral_bounds");
        TCK(dyb_team_propagate(t, md.mode, *t_init, *t_max, tau, save_tau, nullptr), "dyb_team_propagate");
        TCK(dyb_team_get_packets(t, 1, h_PSI_bra, h_PSI_ket), "dyb_team_get_packets");
        return;
    }
    dyb_ctx* c = ctx_for(*n);
    LCK(dyb_upload_hprime(c, h_H, *n), "dyb_upload_hprime");                   // Taylor_gpu.cpp:317
    LCK(dyb_set_packets(c, 1, h_PSI_bra, h_PSI_ket), "dyb_set_packets");
    if (md.bounds)
        LCK(dyb_estimate_spectral_bounds(c, LANCZOS_STEPS, LANCZOS_MARGIN, nullptr, nullptr), "dyb_estimate_spectral_bounds");
    LCK(dyb_propagate(c, md.mode, *t_init, *t_max, tau, save_tau, nullptr), "dyb_propagate");
    LCK(dyb_get_packets(c, 1, h_PSI_bra, h_PSI_ket), "dyb_get_packets");
}

// Taylor_gpu.cpp:743-797, called from diabatic-Ehren.f:115 on the "kernel" MPI rank with H' received from rank 0
void ehrenfestkernel_gpu_(const int* N, const double* h_H, const double* h_A, const double* h_X, double* h_K)
{
    DeviceGuard guard;
    dyb_ctx* c = ctx_for(*N);
    LCK(dyb_upload_hprime(c, h_H, *N), "dyb_upload_hprime");
    LCK(dyb_ehrenfest_kernel(c, h_A, h_X, h_K), "dyb_ehrenfest_kernel");
}

// Taylor_gpu.cpp:801-873 (no caller in the reference tree): the same kernel from the AO packets, the density matrix
// rho / A being formed on the device
void ehrenfestkernel2_gpu_(const int* N, const dyb_complex* h_bra, const dyb_complex* h_ket, const double* h_H, const double* h_X, double* h_K)
{
    DeviceGuard guard;
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
    g_init_by_us = true;
    cudaSetDevice(g_device);
    printf("Process nr. %i using GPU device nr. %i of %i (dynemol_b200)\n", pid ? *pid : 0, g_device, n);
    fflush(stdout);
}

__attribute__((weak)) void gpu_finalize_(void)
{
    if (g_ctx) { dyb_destroy(g_ctx); g_ctx = nullptr; g_N = 0; }
    if (g_team) { dyb_team_destroy(g_team); g_team = nullptr; g_team_N = g_team_P = 0; }
    for (void* p : g_autopinned) { if (cudaHostUnregister(p) != cudaSuccess) cudaGetLastError(); }
    g_autopinned.clear();
}

// The Fortran caller passes n*n*8 as a DEFAULT (32-bit) integer (ElHl_Chebyshev_GPU.f:109-111): it wraps at N = 16384
// (2^31 -> negative) and loses whole multiples of 2^32 from N = 23171 on.  The true size is w + k 2^32 for the smallest k
// that makes size/8 a perfect square (the buffers of this path are N x N doubles).
static size_t unwrap_matrix_bytes(int size_bytes)
{
    const unsigned long long w = (unsigned int)size_bytes;
    for (unsigned long long k = 0; k < 64; ++k) {
        const unsigned long long b = w + (k << 32);
        if (b == 0 || b % 8) continue;
        const unsigned long long e = b / 8, r = (unsigned long long)llround(sqrt((double)e));
        for (unsigned long long q = (r > 1 ? r - 1 : 0); q <= r + 1; ++q) if (q * q == e) return (size_t)b;
    }
    return (size_t)w;
}

// GPU_Interface.cpp:288-302
__attribute__((weak)) void gpu_pin_(void* ptr, int* size_bytes)
{
    cudaError_t e = cudaHostRegister(ptr, unwrap_matrix_bytes(*size_bytes), cudaHostRegisterDefault);
    if (e != cudaSuccess) { printf("ERROR(gpu_pin_): %s\n%s\n\n", cudaGetErrorName(e), cudaGetErrorString(e)); cudaGetLastError(); }
}

__attribute__((weak)) void gpu_unpin_(void* ptr)
{
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { printf("ERROR(gpu_unpin_): %s\n%s\n\n", cudaGetErrorName(e), cudaGetErrorString(e)); cudaGetLastError(); }
}

// el+hole terms (passes over H') of the last propagation made through a legacy symbol (the void symbols cannot return it)
int64_t dyb_legacy_passes_last(void)
{
    if (g_team) return dyb_team_passes_last(g_team);
    if (!g_ctx) return 0;
    int64_t info[16];
    return dyb_get_info(g_ctx, info) == DYB_OK ? info[11] : 0;
}

// host-only: the size recovery of gpu_pin_, exposed for the CPU test-suite (no device needed)
int64_t dyb_unwrap_pin_bytes(int size_bytes) { return (int64_t)unwrap_matrix_bytes(size_bytes); }

}  // extern "C"
