// common.cuh -- shared definitions of the sm_100a propagator kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

namespace dyb {

// ---- geometry of the dual product -----------------------------------------------------------
// A CTA owns a "panel" of PANEL_ROWS matrix rows at a time and sweeps tiles of TILE_COLS columns.
// The panel is split into 8 sub-panels of 256 rows, one per consumer warp; inside a warp, lane l
// owns rows 64*m + 2*l + {0,1}, m = 0..3 (four 128-bit loads per column, 8 rows per thread).
#ifndef DYB_SUB_ROWS
#define DYB_SUB_ROWS 256
#endif
#ifndef DYB_N_CWARPS
#define DYB_N_CWARPS 8
#endif
constexpr int SUB_ROWS     = DYB_SUB_ROWS;         // rows per consumer warp (256: 8 rows per thread; 128: 4 rows per thread)
constexpr int N_CWARPS     = DYB_N_CWARPS;         // consumer warps per CTA (8 with 256 rows, 16 with 128 rows)
constexpr int MPT          = SUB_ROWS / 64;        // 128-bit loads per thread and column (two rows each)
constexpr int PANEL_ROWS   = SUB_ROWS * N_CWARPS;  // 2048
#ifndef DYB_TILE_COLS
#define DYB_TILE_COLS 4
#endif
#ifndef DYB_TMA_STAGES
#define DYB_TMA_STAGES 3
#endif
constexpr int TILE_COLS    = DYB_TILE_COLS;        // columns per tile (TMA box depth): 2 or 4
constexpr int ROW_ALIGN    = SUB_ROWS;             // ld is a multiple of this (zero padded rows)
constexpr int NQ           = 4;                    // reals per index in a "quad" vector: el.re el.im hl.re hl.im
constexpr int TMA_STAGES   = DYB_TMA_STAGES;
constexpr int RED_SLOTS    = 2 * TMA_STAGES;

constexpr int STAGE_H_BYTES  = TILE_COLS * PANEL_ROWS * 8;        // 64 KiB
constexpr int STAGE_X_BYTES  = 128;                               // TILE_COLS*NQ*8 = 64/128 B, padded
constexpr int STAGE_BYTES    = STAGE_H_BYTES + STAGE_X_BYTES;
constexpr int STAGE_TX_BYTES = STAGE_H_BYTES + TILE_COLS * NQ * 8;

// ---- per-term epilogue parameters (host -> kernel, by value) ----------------------------------
struct PartPass {
    int    active;       // particle takes part in this term
    int    k;            // 1-based series index of the term being produced (trace only)
    int    three_term;   // y = alpha*(H x) + beta*x + gamma*x_prev  (Chebyshev);  else y = alpha*(H x)
    int    scale_term;   // term = c*y, else term = y (Taylor: ratio already folded into alpha)
    int    check_conv;   // Convergence(): early exit when both term maxima <= tol and the norm holds
    int    last;         // last term of this series for this particle: latch with ok = norm test (or 0)
    int    last_ok_by_norm; // steady sub-step: ok decided by the norm test alone at the last term
    int    begin;        // chained steady sub-steps (resident kernel): this term starts a new sub-step from the sum of the
                         // previous one (psi <- sum ; sum <- s * psi), Taylor.f:81-126 without the host in the loop
    int    chain;        // this is the last term of a sub-step and another sub-step of the same particle follows in the
                         // same launch: a passed norm test counts the sub-step and does not latch
    int    test_gpu;     // term test of the reference's GPU variant (Taylor_gpu.cpp:84-89,581-587): modulus of the complex element
                         // that holds the largest |re| or |im| (cublasIdamax over 2n reals), compared with a strict `< tol`
    double s_re, s_im;   // scale of the series sum at `begin` (c_0 of the Chebyshev series; 1 for Taylor)
    double alpha_re, alpha_im;
    double beta_re, beta_im;
    double gamma;
    double c_re, c_im;
    double norm_ref;
};

struct PassParams { PartPass part[2]; };

// ---- device-resident series state ------------------------------------------------------------
struct PartState {
    int    latched;      // series over for this particle (converged, failed, or not taking part)
    int    ok;
    int    k_exit;
    int    n_terms;      // terms actually applied in this series
    int    n_sub_ok;     // steady sub-steps (last term without check_conv) that passed the norm test in this launch
    int    pad_;
    double max_b, max_k; // last term maxima
    double dot_re, dot_im;
    double norm;         // | <sum_b | sum_k> | of the last applied term
};

struct Ctrl {
    PartState part[2];
    int       all_latched;
    unsigned  block_counter;      // epilogue "last block" ticket
    int       peer_timeout;       // sticky: a peer's flag did not arrive in time (row-sharded exchange); the host turns it into DYB_ECUDA
    int       pad_;
};

constexpr double TOL_TERM = 1.0e-8;   // Taylor.f:21  (error)
constexpr double TOL_NORM = 1.0e-8;   // Taylor.f:22  (norm_error)

// ---- PTX helpers: mbarrier + TMA (sm_90+/sm_100a) ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// bounded variant for the cooperative one-launch kernels: a copy that never lands must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    for (long long it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > (1ll << 28)) __trap();
}
// 3-D tiled TMA load, global -> shared, completion on an mbarrier, with an L2 cache-policy hint
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both sides 16 B aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with programmaticStreamSerialization may start while its
// predecessor in the stream is still running; it must execute pdl_wait() before touching anything the
// predecessor writes.  pdl_launch_dependents() lets the successor start being scheduled from this point on.
__device__ __forceinline__ void pdl_wait()              { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- cross-GPU flags over NVLink peer memory (system scope) -------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Bounded wait on a peer's epoch flag.  The bound is wall-clock time (%globaltimer, ns), set by the host from
// DYNEMOL_B200_PEER_TIMEOUT_S (default 30 s: peers are launched by independent host threads / processes and may lag by a
// lazy module load or a host stall).  On timeout the kernel does NOT trap (a trap poisons the CUDA context of every rank
// in turn): it raises the sticky Ctrl::peer_timeout flag and returns false; the caller skips its work, later kernels see
// the flag and skip too, and the host returns DYB_ECUDA from the call in progress.
__device__ unsigned long long g_peer_timeout_ns = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool wait_flag_ge(const unsigned long long* p, unsigned long long target, int* timeout_flag) {
    if (ld_acquire_sys(p) >= target) return true;
    const unsigned long long t0 = globaltimer_ns(), limit = g_peer_timeout_ns;
    for (;;) {
        __nanosleep(64);
        if (ld_acquire_sys(p) >= target) return true;
        if (globaltimer_ns() - t0 > limit) { atomicExch(timeout_flag, 1); __threadfence(); return false; }
    }
}

// streaming 128-bit global load that does not allocate in L1 (H' is read exactly once per term)
__device__ __forceinline__ double2 ldg_stream(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

}  // namespace dyb
