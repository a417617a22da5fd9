// matvec.cuh -- the dual product  y_ket = H' x_ket  and  y_bra = H'^T x_bra  in ONE pass over H'.
//
// Replaces the four separate KBLAS-derived GEMV launches per term of the reference
// (dzgemv_kernels.cu:102-495 via Taylor_gpu.cpp:72-90,411-412,570-571) and the MKL dzgemv pair of
// the CPU path (Taylor.f:94-95,186-187; Matrix_math.f:238-301).  Written from scratch for sm_100a.
//
// Data layout:
//   H'   : real, column-major, M rows x Nc cols, leading dimension ld = roundup(M,256), padding rows 0.
//   x,y  : "quad" vectors, 4 doubles per index = (el.re, el.im, hl.re, hl.im), 32 B aligned.
//   ket slabs : [segment][2048 rows][4]   partial H' x_ket over the segment's column range
//   bra slabs : [panel][Ncpad cols][4]    partial H'^T x_bra over the panel's 2048 rows
// The fused epilogue kernel (epilogue.cuh) sums the slabs in a fixed order (deterministic, no atomics).
#pragma once
#include "common.cuh"

namespace dyb {

struct MatvecParams {
    int        M;          // local rows
    int        Nc;         // columns
    long long  ld;         // leading dimension (multiple of 256)
    int        TPP;        // tiles per panel = ceil(Nc / TILE_COLS)
    int        T;          // total tiles = n_panels * TPP (< 2^31 / grid)
    int        Ncpad;      // TPP * TILE_COLS   (bra slab row length)
    const double* H;       // device matrix (LDG variant)
    const double* Xk;      // ket input, quad, indexed by column, padded to Ncpad with zeros
    const double* Xb;      // bra input, quad, indexed by local row, padded to n_panels*2048 with zeros
    double*    ket_slab;
    double*    bra_slab;
    const int* seg_base;   // [grid] first segment index of each CTA
    const Ctrl* ctrl;      // may be null; all_latched => nothing to do
};

// Sum over the 32 lanes of v[0..NV-1] (NV = 8 or 16); on return every lane holds the total of index
// lane >> (NV == 8 ? 2 : 1).  Butterfly with halving (NV/2 + NV/4 + ... exchanged values), then plain
// butterflies for the remaining lane bits.  Fixed order => deterministic.
template <int NV>
__device__ __forceinline__ double transpose_reduce(const double (&v)[NV], int lane) {
    static_assert(NV == 8 || NV == 16, "NV");
    double a[NV / 2];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
            const double keep = hi ? v[NV / 2 + i] : v[i];
            const double send = hi ? v[i] : v[NV / 2 + i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    double b[NV / 4];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
            const double keep = hi ? a[NV / 4 + i] : a[i];
            const double send = hi ? a[i] : a[NV / 4 + i];
            b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    double c[NV / 8];
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < NV / 8; ++i) {
            const double keep = hi ? b[NV / 8 + i] : b[i];
            const double send = hi ? b[i] : b[NV / 8 + i];
            c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    double d;
    if constexpr (NV == 16) {
        const bool hi = lane & 2;
        const double keep = hi ? c[1] : c[0];
        const double send = hi ? c[0] : c[1];
        d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    } else {
        d = c[0] + __shfl_xor_sync(0xffffffffu, c[0], 2);
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}
constexpr int RED_SHIFT = (TILE_COLS * NQ == 16) ? 1 : 2;      // lane -> value index
constexpr int RED_MASK  = (1 << RED_SHIFT) - 1;

// Per-thread state of a consumer: 8 rows (m = 0..3, e = 0..1) x 4 right-hand sides.
struct ConsumerRegs {
    double acc[MPT][2][NQ];   // ket partial sums of the thread's 2*MPT rows
    double xb[MPT][2][NQ];    // bra input at the thread's 2*MPT rows
};

__device__ __forceinline__ void zero_acc(ConsumerRegs& r) {
#pragma unroll
    for (int m = 0; m < MPT; ++m)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int q = 0; q < NQ; ++q) r.acc[m][e][q] = 0.0;
}

// xb for rows  panel*2048 + w*256 + 64m + 2l + e
__device__ __forceinline__ void load_xb(ConsumerRegs& r, const double* __restrict__ Xb, long long panel, int w, int lane) {
    const double2* base = reinterpret_cast<const double2*>(Xb + ((panel * PANEL_ROWS + w * SUB_ROWS + 2 * lane) * NQ));
#pragma unroll
    for (int m = 0; m < MPT; ++m) {
        const double2* p = base + m * (64 * NQ / 2);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double2 v0 = __ldg(p + e * 2), v1 = __ldg(p + e * 2 + 1);
            r.xb[m][e][0] = v0.x; r.xb[m][e][1] = v0.y; r.xb[m][e][2] = v1.x; r.xb[m][e][3] = v1.y;
        }
    }
}

__device__ __forceinline__ void store_acc(const ConsumerRegs& r, double* __restrict__ ket_slab, long long seg, int w, int lane) {
    double2* base = reinterpret_cast<double2*>(ket_slab + ((seg * PANEL_ROWS + w * SUB_ROWS + 2 * lane) * NQ));
#pragma unroll
    for (int m = 0; m < MPT; ++m) {
        double2* p = base + m * (64 * NQ / 2);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            p[e * 2]     = make_double2(r.acc[m][e][0], r.acc[m][e][1]);
            p[e * 2 + 1] = make_double2(r.acc[m][e][2], r.acc[m][e][3]);
        }
    }
}

// One column of the thread's 8 rows: 32 FMAs for the ket, 32 for the bra (two independent chains per
// right-hand side -- even and odd rows -- summed at the end; fixed order, so still deterministic).
__device__ __forceinline__ void fma_column(ConsumerRegs& r, const double2 (&h)[MPT], const double (&xk)[NQ], double (&p)[NQ]) {
    double p1[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        p[q]  = h[0].x * r.xb[0][0][q];
        p1[q] = h[0].y * r.xb[0][1][q];
    }
#pragma unroll
    for (int m = 0; m < MPT; ++m) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            r.acc[m][0][q] = fma(h[m].x, xk[q], r.acc[m][0][q]);
            r.acc[m][1][q] = fma(h[m].y, xk[q], r.acc[m][1][q]);
            if (m > 0) {
                p[q]  = fma(h[m].x, r.xb[m][0][q], p[q]);
                p1[q] = fma(h[m].y, r.xb[m][1][q], p1[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) p[q] += p1[q];
}

// =================================================================================================
// Variant 1: TMA + mbarrier staged, persistent.  8 warps, each a consumer of its own 256-row sub-panel;
// the producer duties rotate: tile t is "retired" (stage released by all 8 warps -> cross-warp sum of its
// bra partials -> TMA refill of the stage with tile t+STAGES) by warp (t mod 8), RETIRE_LAG tiles after
// it was computed.  No dedicated producer warp: with 9 warps one SM sub-partition would host 3 warps and
// cap every thread at 168 registers (16384 per sub-partition); 8 warps leave 255.
// =================================================================================================
constexpr int TMA_THREADS = N_CWARPS * 32;   // 256
#ifndef DYB_RETIRE_LAG
#define DYB_RETIRE_LAG 1
#endif
constexpr int RETIRE_LAG  = DYB_RETIRE_LAG;  // tiles between computing a tile and retiring it

struct TmaSmem {
    // dynamic shared memory carve-up
    static constexpr int off_bar_full  = TMA_STAGES * STAGE_BYTES;
    static constexpr int off_bar_empty = off_bar_full + TMA_STAGES * 8;
    static constexpr int off_red       = off_bar_empty + TMA_STAGES * 8 + 32;   // keep 16 B alignment
    static constexpr int red_bytes     = RED_SLOTS * N_CWARPS * TILE_COLS * NQ * 8;
    static constexpr int total         = ((off_red + red_bytes + 127) / 128) * 128;
};

// tile index inside this CTA -> (panel, column tile)
struct TileCursor {
    int panel, ct;
    __device__ __forceinline__ void next(int TPP) { if (++ct == TPP) { ct = 0; ++panel; } }
};

// A stage is filled by two async copies that complete on the same mbarrier: the H' tile (TMA tensor load) and the
// x_ket values of the tile's columns (bulk copy).  They can be issued at different times: the H' part does not
// depend on the previous kernel's output, the x_ket part does.
__device__ __forceinline__ void issue_tile_h(uint8_t* smem, uint64_t* bar_full, const CUtensorMap* tmap,
                                             int stage, const TileCursor& tc, uint64_t policy) {
    mbar_arrive_expect_tx(&bar_full[stage], STAGE_TX_BYTES);
    tma_load_3d(smem + (size_t)stage * STAGE_BYTES, tmap, &bar_full[stage], 0, tc.panel * N_CWARPS, tc.ct * TILE_COLS, policy);
}
__device__ __forceinline__ void issue_tile_x(uint8_t* smem, uint64_t* bar_full, const double* Xk, int stage, const TileCursor& tc) {
    bulk_load_1d(smem + (size_t)stage * STAGE_BYTES + STAGE_H_BYTES, Xk + (size_t)tc.ct * TILE_COLS * NQ, TILE_COLS * NQ * 8, &bar_full[stage]);
}
__device__ __forceinline__ void issue_tile_x(uint8_t* smem, uint64_t* bar_full, const MatvecParams& P, int stage, const TileCursor& tc) {
    issue_tile_x(smem, bar_full, P.Xk, stage, tc);
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
dual_matvec_tma_kernel(const __grid_constant__ CUtensorMap tmap, const MatvecParams P)
{
    pdl_launch_dependents();        // the epilogue of this term may be queued behind us right away

    // 128 B alignment is all the un-swizzled TMA destinations need; plain pointer arithmetic on the
    // __shared__ array keeps the address space visible to the compiler (LDS/STS, not generic LD/ST)
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar_full  = reinterpret_cast<uint64_t*>(smem + TmaSmem::off_bar_full);
    uint64_t* bar_empty = reinterpret_cast<uint64_t*>(smem + TmaSmem::off_bar_empty);
    double*   red       = reinterpret_cast<double*>(smem + TmaSmem::off_red);

    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b  = blockIdx.x;
    const int t0 = int(((long long)P.T * b) / gridDim.x), t1 = int(((long long)P.T * (b + 1)) / gridDim.x);
    const int nt = t1 - t0;                                          // this CTA's tile range (panel-major)
    if (nt <= 0) return;
    const int panel0 = t0 / P.TPP, ct0 = t0 - panel0 * P.TPP;
    const uint64_t policy = policy_evict_first();                    // H' is streamed once per term

    if (threadIdx.x == 0) {
        prefetch_tensormap(&tmap);
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], N_CWARPS); }
        fence_barrier_init();
        // prologue: start filling the ring with H' tiles -- they do not depend on the previous kernel, so under
        // programmatic dependent launch this overlaps the tail of the preceding epilogue
        TileCursor tc = {panel0, ct0};
        for (int i = 0; i < TMA_STAGES && i < nt; ++i) { issue_tile_h(smem, bar_full, &tmap, i, tc, policy); tc.next(P.TPP); }
    }
    pdl_wait();                     // from here on the previous kernel's vectors / control block are visible
    if (threadIdx.x == 0) {
        TileCursor tc = {panel0, ct0};
        for (int i = 0; i < TMA_STAGES && i < nt; ++i) { issue_tile_x(smem, bar_full, P, i, tc); tc.next(P.TPP); }
    }
    __syncthreads();
    if (P.ctrl != nullptr && P.ctrl->all_latched) {                // series already decided: skip the pass,
        if (threadIdx.x == 0)                                      // but never exit with copies in flight to our smem
            for (int i = 0; i < TMA_STAGES && i < nt; ++i) mbar_wait(&bar_full[i], 0u);
        return;
    }

    ConsumerRegs r;
    zero_acc(r);
    int seg = P.seg_base[b];
    TileCursor cur = {panel0, ct0};        // tile j being computed
    TileCursor ret = {panel0, ct0};        // tile jr = j - RETIRE_LAG being retired
    TileCursor nxt = {panel0, ct0};        // tile jr + STAGES being issued at retirement
    for (int i = 0; i < TMA_STAGES; ++i) nxt.next(P.TPP);
    int s = 0, ph = 0, rslot = 0;          // ring position / phase / red slot of tile j
    int sr = 0, phr = 0, rslotr = 0;       // the same for tile jr

    for (int j = 0; j < nt + RETIRE_LAG; ++j) {
        if (j < nt) {
            if (j == 0 || cur.ct == 0) load_xb(r, P.Xb, cur.panel, w, lane);

            mbar_wait(&bar_full[s], uint32_t(ph));
            const uint8_t* stage = smem + (size_t)s * STAGE_BYTES;
            const double*  sH = reinterpret_cast<const double*>(stage);
            const double2* sX = reinterpret_cast<const double2*>(stage + STAGE_H_BYTES);

            double pv[TILE_COLS * NQ];
#ifdef DYB_NO_MATH      // ceiling experiment: stream the tiles, touch one value per stage, do no arithmetic
#pragma unroll
            for (int q = 0; q < TILE_COLS * NQ; ++q) pv[q] = 0.0;
            r.acc[0][0][0] += sH[(size_t)w * SUB_ROWS + lane] + sX[0].x;
#else
#pragma unroll
            for (int c = 0; c < TILE_COLS; ++c) {
                const double2 x01 = sX[c * 2], x23 = sX[c * 2 + 1];
                const double xk[NQ] = {x01.x, x01.y, x23.x, x23.y};
                const double2* hp = reinterpret_cast<const double2*>(sH + (size_t)(c * N_CWARPS + w) * SUB_ROWS) + lane;
                double2 h[MPT];
#pragma unroll
                for (int m = 0; m < MPT; ++m) h[m] = hp[m * 32];
                double p[NQ];
                fma_column(r, h, xk, p);
#pragma unroll
                for (int q = 0; q < NQ; ++q) pv[c * NQ + q] = p[q];
            }
#endif
            const double tot = transpose_reduce<TILE_COLS * NQ>(pv, lane);
            if ((lane & RED_MASK) == 0)
                red[rslot * (N_CWARPS * TILE_COLS * NQ) + w * (TILE_COLS * NQ) + (lane >> RED_SHIFT)] = tot;
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[s]);      // release: stage reads and red[] writes are done

            if (j == nt - 1 || cur.ct == P.TPP - 1) {       // end of this CTA's segment of the panel
                store_acc(r, P.ket_slab, seg, w, lane);
                zero_acc(r);
                ++seg;
            }
            cur.next(P.TPP);
            if (++s == TMA_STAGES) { s = 0; ph ^= 1; }
            if (++rslot == RED_SLOTS) rslot = 0;
        }
        if (j >= RETIRE_LAG) {
            const int jr = j - RETIRE_LAG;
            if ((jr & (N_CWARPS - 1)) == w) {
                // all 8 warps have released the stage of tile jr (and published its bra partials)
                mbar_wait(&bar_empty[sr], uint32_t(phr));
                if (lane < TILE_COLS * NQ) {
                    const double* rs = red + rslotr * (N_CWARPS * TILE_COLS * NQ) + lane;
                    double sum = rs[0];
#pragma unroll
                    for (int ww = 1; ww < N_CWARPS; ++ww) sum += rs[ww * TILE_COLS * NQ];
                    P.bra_slab[((size_t)ret.panel * P.Ncpad + (size_t)ret.ct * TILE_COLS) * NQ + lane] = sum;
                }
                if (lane == 0 && jr + TMA_STAGES < nt) { issue_tile_h(smem, bar_full, &tmap, sr, nxt, policy); issue_tile_x(smem, bar_full, P, sr, nxt); }
                __syncwarp();
            }
            ret.next(P.TPP); nxt.next(P.TPP);
            if (++sr == TMA_STAGES) { sr = 0; phr ^= 1; }
            if (++rslotr == RED_SLOTS) rslotr = 0;
        }
    }
}

// =================================================================================================
// Variant 2: direct 128-bit streaming global loads (no shared-memory staging of H').  Same tiling,
// same outputs; kept as the baseline the TMA pipeline is measured against and as a cross-check.
// =================================================================================================
constexpr int LDG_THREADS = N_CWARPS * 32;   // 256

__device__ __forceinline__ void ldg_tile(double2 (&h)[TILE_COLS][MPT], const MatvecParams& P, int panel, int ct, int w, int lane) {
    const long long rb = (long long)panel * PANEL_ROWS + (long long)w * SUB_ROWS;   // first row of this warp's sub-panel
    const bool rows_ok = rb < P.ld;                                           // ld is a multiple of 256
#pragma unroll
    for (int c = 0; c < TILE_COLS; ++c) {
        const long long col = (long long)ct * TILE_COLS + c;
        const bool ok = rows_ok && col < P.Nc;
        const double* src = P.H + col * P.ld + rb + 2 * lane;
#pragma unroll
        for (int m = 0; m < MPT; ++m) h[c][m] = ok ? ldg_stream(src + m * 64) : make_double2(0.0, 0.0);
    }
}

__global__ void __launch_bounds__(LDG_THREADS, 1)
dual_matvec_ldg_kernel(const MatvecParams P)
{
    pdl_launch_dependents();
    pdl_wait();
    if (P.ctrl != nullptr && P.ctrl->all_latched) return;
    __shared__ double red[2][N_CWARPS][TILE_COLS * NQ];

    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b  = blockIdx.x;
    const int t0 = int(((long long)P.T * b) / gridDim.x), t1 = int(((long long)P.T * (b + 1)) / gridDim.x);
    const int nt = t1 - t0;
    if (nt <= 0) return;

    ConsumerRegs r;
    zero_acc(r);
    int seg = P.seg_base[b];
    int panel = t0 / P.TPP, ct = t0 - panel * P.TPP;

    double2 hbuf[TILE_COLS][MPT];
    ldg_tile(hbuf, P, panel, ct, w, lane);

    for (int j = 0; j < nt; ++j) {
        if (j == 0 || ct == 0) load_xb(r, P.Xb, panel, w, lane);

        double2 h[TILE_COLS][MPT];
#pragma unroll
        for (int c = 0; c < TILE_COLS; ++c)
#pragma unroll
            for (int m = 0; m < MPT; ++m) h[c][m] = hbuf[c][m];
        if (j + 1 < nt) {                                                                 // prefetch next tile
            const bool wrap = (ct + 1 == P.TPP);
            ldg_tile(hbuf, P, wrap ? panel + 1 : panel, wrap ? 0 : ct + 1, w, lane);
        }

        double pv[TILE_COLS * NQ];
#pragma unroll
        for (int c = 0; c < TILE_COLS; ++c) {
            const double2* xp = reinterpret_cast<const double2*>(P.Xk + ((size_t)ct * TILE_COLS + c) * NQ);
            const double2 x01 = __ldg(xp), x23 = __ldg(xp + 1);
            const double xk[NQ] = {x01.x, x01.y, x23.x, x23.y};
            double p[NQ];
            fma_column(r, h[c], xk, p);
#pragma unroll
            for (int q = 0; q < NQ; ++q) pv[c * NQ + q] = p[q];
        }
        const double tot = transpose_reduce<TILE_COLS * NQ>(pv, lane);
        if ((lane & RED_MASK) == 0) red[j & 1][w][lane >> RED_SHIFT] = tot;
        __syncthreads();
        if (threadIdx.x < TILE_COLS * NQ) {
            double sum = red[j & 1][0][threadIdx.x];
#pragma unroll
            for (int ww = 1; ww < N_CWARPS; ++ww) sum += red[j & 1][ww][threadIdx.x];
            P.bra_slab[((size_t)panel * P.Ncpad + (size_t)ct * TILE_COLS) * NQ + threadIdx.x] = sum;
        }
        if (j == nt - 1 || ct == P.TPP - 1) {
            store_acc(r, P.ket_slab, seg, w, lane);
            zero_acc(r);
            ++seg;
        }
        if (++ct == P.TPP) { ct = 0; ++panel; }
    }
}

}  // namespace dyb
