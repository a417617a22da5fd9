// mid.cuh -- mid-size operators (roughly 1800 < N < 7000): ONE cooperative launch per series, H' STREAMED per term,
// NO grid barrier: the CTAs talk through epoch-tagged words in the L2.
//
// Between the shared-memory-resident kernel (resident.cuh, N <= 1824) and the bandwidth regime of the two-launch path
// (matvec.cuh + epilogue.cuh, N >~ 8000) a term is neither: H' (27 ... 300 MB) sits in or near the 126 MB L2, one pass over
// it takes 5 ... 40 us, and the two-launch path adds a nearly constant ~22 us on top (launch pair, pipeline fill, 74-148 ket
// partials per row in the epilogue, last-block decision; profiles/midrange_r1.md).  This kernel replaces, for that range,
// the per-term launches of the reference (four KBLAS GEMVs + scal + axpy + 2 Idamax + dotc and three host round trips per
// term, Taylor_gpu.cpp:570-600, dzgemv_kernels.cu:102-495) by
//
//   * a Gr x Gc grid of CTAs (<= one per SM); CTA (bi, bj) owns the block rows [bi*R, bi*R+R) x cols [bj*Cn, bj*Cn+Cn) of
//     H' for the whole series.  R = 256*WR rows, Cn = a multiple of the tile width;
//   * a TMA ring (cp.async.bulk.tensor.3d, mbarrier full/empty pairs) that streams the block as tiles of R rows x TC
//     columns (32 KiB) and RUNS ACROSS THE TERMS: H' does not change, so the first tiles of term t+1 are in flight while
//     the exchange of term t runs;
//   * the tile engine of the dual product (matvec.cuh): 8 rows x 2 columns per thread and tile, ket sums in registers,
//     bra sums by a lane butterfly (transpose_reduce) that is software-pipelined one tile behind the FMAs -- one pass
//     serves H'x_ket and H'^T x_bra of electron and hole;
//   * an OWNER-REDUCE exchange without any grid-wide barrier.  Every index g has one owner CTA (g / E).  Per term a CTA
//       1. publishes its partial products (R ket entries of its block row, Cn bra entries of its block column),
//       2. as OWNER: collects the Gc ket partials and the Gr bra partials of its E indices, applies the recurrence and
//          the series sum to ITS copy of the state (x, x_prev, sum, start vector: shared memory, 2 x E entries), and
//          publishes the new vector entries and its 8 convergence scalars,
//       3. as CONSUMER: collects the R + Cn new entries it multiplies next.
//     Every published double travels as a 16-byte word {lo32, epoch, hi32, epoch} written and read with ONE 128-bit
//     relaxed access whose two 64-bit halves are single-copy atomic (the LL protocol of NCCL): a reader polls the word
//     itself until both tags carry the epoch of the term -- no fence, no flag, no counter.  Epochs grow across launches,
//     buffers are double-buffered by term parity (scalars: four deep); the data dependencies of the recurrence guarantee
//     that a slot is never overwritten before its readers are done (a CTA needs x(t+2) from exactly the owners that read
//     its partials of term t; checked on the host for the plans in use: tests/test_host_logic.py,
//     test_mid_exchange_needs_no_barrier);
//   * the decision on term t-1 (decide_particle, the code of the other two paths) is taken by every CTA, identically,
//     from the owners' scalars before term t is applied: a latched particle skips the update.
//
// Same PassParams / Ctrl contract as resident.cuh: propagate_series does not know which kernel ran.  Chained steady
// sub-steps (PartPass::begin / chain) are supported; the reference-GPU term test (PartPass::test_gpu) is not -- those parity
// modes stay on the two-launch path.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"
#include "matvec.cuh"

namespace dyb {

constexpr int MID_THREADS     = 256;
constexpr int MID_WARPS       = 8;
constexpr int MID_SUB         = 256;        // rows per warp (8 per thread: lane l owns rows 64m + 2l + {0,1}, m = 0..3)
constexpr int MID_MPT         = 4;
constexpr int MID_CPW         = 2;          // columns per warp and tile
constexpr int MID_STAGE_BYTES = 32768;      // R x TC x 8 B with R*TC = 4096 for every WR
constexpr int MID_MAX_ST      = 5;
constexpr int MID_U_BYTES     = 32768;      // union region: bra partials of the row groups / ket tree reduction
constexpr int MID_MAX_E       = 64;         // indices an owner may hold
#ifndef DYB_MID_CWP
#define DYB_MID_CWP 17
#endif
constexpr int MID_CWP         = DYB_MID_CWP; // owner: partial words per thread ((Gr + Gc) * E * 4 <= 256 * 17) ...
constexpr int MID_CWS         = 5;          // ... and scalar words per thread (8 * owners <= 256 * 5), all in flight at once
constexpr int MID_XT          = 224;        // consumer threads: warps 1..7 (warp 0 publishes the scalars meanwhile)
constexpr int MID_XU          = 19;         // consumer: words per thread ((R + Cn) * 4 <= 224 * 19)
constexpr int MID_SCU         = 10;         // decision: owners per thread (16 * 10 >= owner CTAs)
constexpr int MID_SMEM_MAX    = 227 * 1024 - 2048;
constexpr int MID_NO_WORD     = 0xffff;     // table entries of a word that does not exist (index >= N): 16-bit / 32-bit table
constexpr int MID_NO_WORD32   = INT_MIN;
#ifndef DYB_LL_SLEEP
#define DYB_LL_SLEEP 300                    // ns between two polling rounds of a thread
#endif
constexpr long long MID_SPIN_LIMIT = 1ll << 22;     // polls before a reader gives up (a word that never arrives must end in a
                                                    // trap, not in a hung GPU)

struct MidParams {
    int N, Gr, Gc, Cnp, NT, ST;             // grid, block width (multiple of TC), tiles per term, ring depth
    int E, n_own;                           // indices per owner CTA (ceil(N / grid)), CTAs that own at least one index
    int tab16;                              // the collect table holds 16-bit (source, word) pairs instead of 32-bit offsets (large blocks)
    unsigned epoch0;                        // the epoch of this launch's term t is epoch0 + t + 1
    float l2_frac;                          // fraction of the H' lines loaded with L2::evict_last (0: plain evict_first stream)
    const double* x0k; const double* x0b;   // starting vectors (quads), written by series_init_kernel
    double* sum_b; double* sum_k;           // in: series sums at the start; out: at the latch / end of the series
    ulonglong2* pk; ulonglong2* pb;         // [2][Gr][Gc][R][NQ] / [2][Gc][Gr][Cnp][NQ] tagged partial products (term parity first)
    ulonglong2* xx;                         // [2][2][N][NQ] tagged new vector entries (parity, side: 0 ket / 1 bra)
    ulonglong2* sc;                         // [4][grid][8] tagged scalars of the owners (term mod 4)
    Ctrl* ctrl;
    const PassParams* passes; int n_steps;
    long long* prof;                        // DYB_SERIES_PROF builds: [32][grid][8] clock64 stamps (else null)
};

#ifdef DYB_SERIES_PROF
#define DYB_MSTAMP(i) do { if (threadIdx.x == 0 && t < 32) P.prof[((size_t)t * G + blockIdx.x) * 16 + (i)] = clock64(); } while (0)
#define DYB_MSTAMP_T(i, thr, tt) do { if (threadIdx.x == (thr) && (tt) < 32 && (tt) >= 0) P.prof[((size_t)(tt) * G + blockIdx.x) * 16 + (i)] = clock64(); } while (0)
#define DYB_MROUNDS(i, v) do { if (threadIdx.x == 0 && t < 32) P.prof[((size_t)t * G + blockIdx.x) * 16 + (i)] = (v); } while (0)
#else
#define DYB_MSTAMP(i) do { } while (0)
#define DYB_MSTAMP_T(i, thr, tt) do { } while (0)
#define DYB_MROUNDS(i, v) do { } while (0)
#endif

struct MidSmem {                            // dynamic shared memory carve-up (byte offsets)
    int bars, U, xk, xb, own, tab, total;
    __host__ __device__ MidSmem(int ST, int R, int Cnp, int E, int Gsum, int n_own, int tab16) {     // Gsum = Gr + Gc
        bars = ST * MID_STAGE_BYTES;
        U    = bars + 128;                  // union region; also holds what an owner collects per term: (Gr + Gc) * E * 4
        const int ub = (Gsum * E * NQ + n_own * 8) * 8;        // partials and the 8 scalars of every owner
        xk   = U + (ub > MID_U_BYTES ? (ub + 127) & ~127 : MID_U_BYTES);
        xb   = xk + Cnp * 32;               // [Cnp][NQ] then [R][NQ]: contiguous (the consumer fills both as one array)
        own  = xb + R * 32;                 // owner state: cur, prev, sum, start [2 sides][E][NQ] + magnitudes [2][E][2]
        tab  = own + E * (4 * 2 * NQ * 8 + 2 * 2 * 8);
        // the owner's collect table, [slot][thread]: the 32-bit offset of every word, or (when that would cost a ring stage)
        // 16-bit (source, e * 4 + q) pairs behind the offsets per source [Gsum] and per (side, e, q) [2][E * 4]
        const int slots = (Gsum * E * NQ + MID_THREADS - 1) / MID_THREADS;
        total = tab + (tab16 ? ((Gsum + 2 * E * NQ) * 4 + 127) / 128 * 128 + slots * MID_THREADS * 2 : slots * MID_THREADS * 4);
    }
};

// ---- epoch-tagged words: one double per 16 B, {lo32 | epoch << 32, hi32 | epoch << 32}
__device__ __forceinline__ void ll_store(ulonglong2* p, double v, unsigned ep) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), e = (unsigned long long)ep << 32;
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"((b & 0xffffffffull) | e), "l"((b >> 32) | e) : "memory");
}
__device__ __forceinline__ ulonglong2 ll_load(const ulonglong2* p) {
    ulonglong2 r;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void ll_load2(const ulonglong2* p, unsigned long long& x, unsigned long long& y) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "l"(p) : "memory");
}
__device__ __forceinline__ bool ll_valid(const ulonglong2& r, unsigned ep) {
    return (unsigned)(r.x >> 32) == ep && (unsigned)(r.y >> 32) == ep;
}
__device__ __forceinline__ double ll_value(const ulonglong2& r) {
    return __longlong_as_double((long long)((r.x & 0xffffffffull) | (r.y << 32)));
}
// Batched polling.  The words whose bit is set in `pend` (bit i: address ADDR(i)) are requested TOGETHER, tested, and the
// missing ones requested again together -- a round costs one L2 round trip whatever the number of words; OUT(i, value) consumes
// word i as soon as it is valid.  A macro on purpose: the 16-byte words must live in REGISTERS (an array that ends up in local
// memory -- lambdas / references did that -- puts a store behind every load and serialises them, one round trip each).
#define DYB_LL_POLL(NW, pend, EP, ADDR, OUT, AFTER_REQ)                                                                    \
    do {                                                                                                          \
        for (long long it_ = 0;; ++it_) {                                                                         \
            unsigned long long rx_[NW], ry_[NW];                                                                  \
            _Pragma("unroll") for (int i = 0; i < (NW); ++i) {                                                    \
                rx_[i] = 0ull; ry_[i] = 0ull;                                                                     \
                if (((pend) >> i) & 1u) ll_load2(ADDR(i), rx_[i], ry_[i]);                                        \
            }                                                                                                     \
            if (it_ == 0) { AFTER_REQ; }                                                                          \
            _Pragma("unroll") for (int i = 0; i < (NW); ++i)                                                      \
                if ((((pend) >> i) & 1u) && (unsigned)(rx_[i] >> 32) == EP(i) && (unsigned)(ry_[i] >> 32) == EP(i)) {   \
                    OUT(i, __longlong_as_double((long long)((rx_[i] & 0xffffffffull) | (ry_[i] << 32))));         \
                    (pend) &= ~(1u << i);                                                                         \
                }                                                                                                 \
            if (!(pend)) { ll_rounds = (int)it_ + 1; break; }                                                     \
            if (it_ > MID_SPIN_LIMIT) __trap();                                                                   \
            if (DYB_LL_SLEEP > 0) __nanosleep(DYB_LL_SLEEP);                                                      \
        }                                                                                                         \
    } while (0)

// The same for two word lists polled together (the owner's partials and the owners' scalars: different epochs and addressing).
#define DYB_LL_POLL2(NA, pa, EPA, ADDRA, OUTA, NB, pb, EPB, ADDRB, OUTB, AFTER_REQ)                               \
    do {                                                                                                          \
        for (long long it_ = 0;; ++it_) {                                                                         \
            unsigned long long ax_[NA], ay_[NA], bx_[NB], by_[NB];                                                \
            _Pragma("unroll") for (int i = 0; i < (NA); ++i) {                                                    \
                ax_[i] = 0ull; ay_[i] = 0ull;                                                                     \
                if (((pa) >> i) & 1u) ll_load2(ADDRA(i), ax_[i], ay_[i]);                                         \
            }                                                                                                     \
            _Pragma("unroll") for (int i = 0; i < (NB); ++i) {                                                    \
                bx_[i] = 0ull; by_[i] = 0ull;                                                                     \
                if (((pb) >> i) & 1u) ll_load2(ADDRB(i), bx_[i], by_[i]);                                         \
            }                                                                                                     \
            if (it_ == 0) { AFTER_REQ; }                                                                          \
            _Pragma("unroll") for (int i = 0; i < (NA); ++i)                                                      \
                if ((((pa) >> i) & 1u) && (unsigned)(ax_[i] >> 32) == (EPA) && (unsigned)(ay_[i] >> 32) == (EPA)) { \
                    OUTA(i, __longlong_as_double((long long)((ax_[i] & 0xffffffffull) | (ay_[i] << 32))));        \
                    (pa) &= ~(1u << i);                                                                           \
                }                                                                                                 \
            _Pragma("unroll") for (int i = 0; i < (NB); ++i)                                                      \
                if ((((pb) >> i) & 1u) && (unsigned)(bx_[i] >> 32) == (EPB) && (unsigned)(by_[i] >> 32) == (EPB)) { \
                    OUTB(i, __longlong_as_double((long long)((bx_[i] & 0xffffffffull) | (by_[i] << 32))));        \
                    (pb) &= ~(1u << i);                                                                           \
                }                                                                                                 \
            if (!((pa) | (pb))) { ll_rounds = (int)it_ + 1; break; }                                              \
            if (it_ > MID_SPIN_LIMIT) __trap();                                                                   \
            if (DYB_LL_SLEEP > 0) __nanosleep(DYB_LL_SLEEP);                                                      \
        }                                                                                                         \
    } while (0)

// 32 FMAs of half a column: rows of m = M0, M0 + 1 against the 4 reals of x_ket (-> acc) and of x_bra (-> p, p1)
template <int M0>
__device__ __forceinline__ void mid_fma_half(double (&acc)[MID_MPT][2][NQ], const double (&xb)[MID_MPT][2][NQ],
                                             const double2 (&h)[MID_MPT], const double (&xk)[NQ], double (&p)[NQ], double (&p1)[NQ]) {
#pragma unroll
    for (int m = M0; m < M0 + 2; ++m) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            acc[m][0][q] = fma(h[m].x, xk[q], acc[m][0][q]);
            acc[m][1][q] = fma(h[m].y, xk[q], acc[m][1][q]);
            if (m == 0) { p[q] = h[m].x * xb[m][0][q]; p1[q] = h[m].y * xb[m][1][q]; }
            else        { p[q] = fma(h[m].x, xb[m][0][q], p[q]); p1[q] = fma(h[m].y, xb[m][1][q], p1[q]); }
        }
    }
}

// static shared state of a CTA (one instance in the kernel; the exchange functions get a pointer)
struct MidShared {
    MidParams  P;                           // copy of the launch parameters: the non-inlined functions read them from shared
                                            // memory (through a reference to the kernel parameter every read was a generic
                                            // load from parameter space: 4000 cycles before the first request of a term)
    Ctrl       sctrl;
    PassParams spass[2];
    double     wsc[4][8];                   // mid_decide (last term): per-warp combinations of the scalars
    double     wsc8[8][8];                  // mid_owner: the same, all eight warps
    int        stop_chain;                  // a particle failed a chained sub-step: the other one stops at its next sub-step
};                                          // boundary so that both resume together

// The exchange of a term lives in functions that are NOT inlined into the kernel: the tile engine keeps ~230 registers busy,
// and inlined next to it ptxas parked the 16-byte words of the polling loops in local memory (a store behind every load =
// the loads of a thread serialised, one L2 round trip each: 10 000 cycles to request 16 words).  On their own these functions
// need few registers and every polling round is one batch of independent loads.

// Decision on term td: scalars of all owners, fixed combination order, identical in every CTA (decide_particle = the code
// of the two other paths, Taylor.f:194-207 / :102-105).
__device__ __noinline__ void mid_decide(MidShared* sh, int td, bool allow_chain) {
    const MidParams& P = sh->P;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int G = P.Gr * P.Gc;
    if (tid >= 128) {
        const int j = tid - 128, s = j & 7, ob = j >> 3;
        const unsigned epd = P.epoch0 + td + 1;
        const ulonglong2* src = P.sc + ((size_t)(td & 3) * G) * 8 + s;
        unsigned need = 0;
#pragma unroll
        for (int i = 0; i < MID_SCU; ++i) if (ob + 16 * i < P.n_own) need |= 1u << i;
        double sv[MID_SCU];
        [[maybe_unused]] int ll_rounds = 0;
#define DYB_SC_ADDR(i) (src + (size_t)(ob + 16 * (i)) * 8)
#define DYB_SC_OUT(i, v) sv[i] = (v)
        unsigned pend = need;
#define DYB_SC_EP(i) epd
        DYB_LL_POLL(MID_SCU, pend, DYB_SC_EP, DYB_SC_ADDR, DYB_SC_OUT, (void)0);
#undef DYB_SC_EP
#undef DYB_SC_ADDR
#undef DYB_SC_OUT
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < MID_SCU; ++i)
            if ((need >> i) & 1u) a = ((s & 3) < 2) ? fmax(a, sv[i]) : a + sv[i];
#pragma unroll
        for (int off = 8; off < 32; off <<= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, a, off);
            a = ((s & 3) < 2) ? fmax(a, o) : a + o;
        }
        if (lane < 8) sh->wsc[w - 4][lane] = a;
    }
    __syncthreads();
    if (tid == 0 || tid == 32) {
        const int p = tid >> 5;
        double fin[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int s = 4 * p + q;
            fin[q] = (q < 2) ? fmax(fmax(sh->wsc[0][s], sh->wsc[1][s]), fmax(sh->wsc[2][s], sh->wsc[3][s]))
                             : (sh->wsc[0][s] + sh->wsc[1][s]) + (sh->wsc[2][s] + sh->wsc[3][s]);
        }
        decide_particle(sh->sctrl.part[p], sh->spass[td & 1].part[p], fin, allow_chain);
    }
    __syncthreads();
}

// Owner side of a term, one function:
//   1. ONE polling batch collects the partials of the owned indices (term t) AND the 8 scalars of every owner (term t-1)
//      into val[] (shared memory).  Word w < W: w = k * E4 + e * 4 + q -- source k (0 .. Gc-1: ket partial of block column k;
//      Gc .. Gc+Gr-1: bra partial of block row k - Gc), owned index e, real q; word W + o * 8 + s: scalar s of owner o.
//      Thread tid takes the words tid + 256 j (<= 17 partial words, <= 5 scalar words), all in flight at once.
//   2. warp 0 takes the decision on term t-1 (decide_particle = the code of the two other paths, Taylor.f:194-207 / :102-105;
//      fixed combination order, identical in every CTA) while warps 1..7 add up the partials (fixed order);
//   3. warps 1..7 apply the recurrence and the series sum to the owner state (own: cur, prev, sum, start [2 sides][E][NQ] +
//      magnitudes [2][E][2]) and publish the new entries; a latched particle skips the update;
//   4. warp 0 publishes the convergence scalars of the owned indices while warps 1..7 already collect their next input.
// Returns true when both particles are latched (the series is over; nothing of term t was applied).
template <int R>
__device__ __noinline__ bool mid_owner(MidShared* sh, double* val, double* own, int t) {
    const MidParams& P = sh->P;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int Gr = P.Gr, Gc = P.Gc, Cnp = P.Cnp, N = P.N, E = P.E, G = Gr * Gc;
    const int E4 = E * NQ, W = (Gc + Gr) * E4, o0 = blockIdx.x * E;
    const int par = t & 1;
    const unsigned ep = P.epoch0 + t + 1;
    {
        // offsets of the thread's partial words, from the tables built once per launch: wtab32[slot][thread] = offset (ket:
        // >= 0 from pkb, bra: ~offset from pbb), or (tab16) wtab16 = (k << 8 | e * 4 + q) with kt[k] = offset of source k,
        // rt[side][e * 4 + q] = offset of the word inside a source
        const int* kt = reinterpret_cast<const int*>(own + 4 * 2 * E * NQ + 2 * E * 2);
        int off[MID_CWP];
        unsigned pp = 0, ps = 0;                                 // pending partial / scalar words
        if (P.tab16) {
            const int* rt = kt + (Gc + Gr);
            const unsigned short* wtab16 = reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(kt) + ((Gc + Gr + 2 * E4) * 4 + 127) / 128 * 128) + tid;
#pragma unroll
            for (int j = 0; j < MID_CWP; ++j) {
                off[j] = 0;
                if (j * MID_THREADS < W) {                       // uniform
                    const int kr = wtab16[j * MID_THREADS];
                    if (kr != MID_NO_WORD) {
                        pp |= 1u << j;
                        const int k = kr >> 8, isb = k >= Gc;
                        const int o = kt[k] + rt[(isb ? E4 : 0) + (kr & 255)];
                        off[j] = isb ? ~o : o;
                    }
                }
            }
        } else {
            const int* wtab32 = kt + tid;
#pragma unroll
            for (int j = 0; j < MID_CWP; ++j) {
                off[j] = MID_NO_WORD32;
                if (j * MID_THREADS < W) off[j] = wtab32[j * MID_THREADS];          // uniform guard; the table is padded to whole slots
                if (off[j] != MID_NO_WORD32) pp |= 1u << j;
            }
        }
        const int nS = t > 0 ? P.n_own * 8 : 0;                  // scalars of term t-1: word o * 8 + s
#pragma unroll
        for (int j = 0; j < MID_CWS; ++j) if (j * MID_THREADS + tid < nS) ps |= 1u << j;
        const ulonglong2* pkb = P.pk + (size_t)par * Gr * Gc * R * NQ;
        const ulonglong2* pbb = P.pb + (size_t)par * Gc * Gr * Cnp * NQ;
        const ulonglong2* scb = P.sc + (size_t)((t + 3) & 3) * G * 8 + tid;
        [[maybe_unused]] int ll_rounds = 0;
#define DYB_CO_ADDR(j) (off[j] >= 0 ? pkb + off[j] : pbb + ~off[j])
#define DYB_CO_OUT(j, v) val[(j) * MID_THREADS + tid] = (v)
#define DYB_CS_ADDR(j) (scb + (j) * MID_THREADS)
#define DYB_CS_OUT(j, v) val[W + (j) * MID_THREADS + tid] = (v)
        DYB_MSTAMP(9);
        // (the ket rows staged in U are overwritten by val: every thread must have published them -> barrier behind the request)
        DYB_LL_POLL2(MID_CWP, pp, ep, DYB_CO_ADDR, DYB_CO_OUT, MID_CWS, ps, ep - 1u, DYB_CS_ADDR, DYB_CS_OUT, __syncthreads());
        DYB_MROUNDS(8, ll_rounds);
#undef DYB_CO_ADDR
#undef DYB_CO_OUT
#undef DYB_CS_ADDR
#undef DYB_CS_OUT
    }
    DYB_MSTAMP(7);
    __syncthreads();
    DYB_MSTAMP(10);

    // ---- decision on term t-1, first half (all warps): thread (warp w, lane) = (slot s = lane & 7, sub = lane >> 3) combines the
    // owners w * 4 + sub + 32 i of its slot, the four lanes of a slot are joined by two shuffles: 8 values per warp.  Done by
    // one warp alone this was 1700 cycles per term.  The maxima are non-negative and never NaN (the owners start from 0.0 and
    // fmax drops NaNs), so their order is the order of their bit patterns: integer max.  Fixed order, the same in every CTA.
    if (t > 0) {
        const int s = lane & 7, ob = w * 4 + (lane >> 3), n_own = P.n_own;
        const bool is_max = (s & 3) < 2;
        const double* sv = val + W + s;
        double v[MID_CWS];
#pragma unroll
        for (int i = 0; i < MID_CWS; ++i) v[i] = (ob + 32 * i < n_own) ? sv[(size_t)(ob + 32 * i) * 8] : 0.0;
        double a;
        if (is_max) {
            long long m = __double_as_longlong(v[0]);
#pragma unroll
            for (int i = 1; i < MID_CWS; ++i) m = max(m, __double_as_longlong(v[i]));
            a = __longlong_as_double(m);
        } else {
            a = ((v[0] + v[1]) + (v[2] + v[3])) + v[4];
            static_assert(MID_CWS == 5, "combination order of the scalars");
        }
#pragma unroll
        for (int off = 8; off < 32; off <<= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, a, off);
            a = is_max ? __longlong_as_double(max(__double_as_longlong(a), __double_as_longlong(ov))) : a + ov;
        }
        if (lane < 8) sh->wsc8[w][lane] = a;
    }
    __syncthreads();

    // ---- warp 0: decision on term t-1; warps 1..7: sums of the partials.  task = (side, e, q), MID_XT threads, <= 3 each;
    // the two reals of a complex value sit in neighbouring lanes
    double* ocur = own;
    double* oprv = ocur + 2 * E * NQ;
    double* osum = oprv + 2 * E * NQ;
    double* opsi = osum + 2 * E * NQ;
    double* omag = opsi + 2 * E * NQ;
    double s_old[3] = {0.0, 0.0, 0.0}, s_cur[3] = {0.0, 0.0, 0.0}, s_prv[3] = {0.0, 0.0, 0.0}, s_sum[3] = {0.0, 0.0, 0.0},
           s_psi[3] = {0.0, 0.0, 0.0}, s_mag[3] = {0.0, 0.0, 0.0};
    unsigned flags = 0;                                           // per round u: bit 2u = update computed, bit 2u+1 = sub-step begins
    if (w == 0) {
        if (t > 0) {
            // second half: lane s < 8 joins the eight warps' values of slot s (a fixed tree)
            const int s = lane & 7;
            const bool is_max = (s & 3) < 2;
            double c[8];
#pragma unroll
            for (int ww = 0; ww < 8; ++ww) c[ww] = sh->wsc8[ww][s];
            double a;
            if (is_max) {
                long long m = __double_as_longlong(c[0]);
#pragma unroll
                for (int ww = 1; ww < 8; ++ww) m = max(m, __double_as_longlong(c[ww]));
                a = __longlong_as_double(m);
            } else a = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
            double fin[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) fin[q] = __shfl_sync(0xffffffffu, a, 4 * (lane & 1) + q);
            DYB_MSTAMP(12);
            // stop_chain as of the previous term: the same in every CTA
            if (lane < 2) decide_particle(sh->sctrl.part[lane], sh->spass[(t - 1) & 1].part[lane], fin, !sh->stop_chain);
            __syncwarp();
            if (lane == 0 && ((sh->sctrl.part[0].latched && !sh->sctrl.part[0].ok) || (sh->sctrl.part[1].latched && !sh->sctrl.part[1].ok))) sh->stop_chain = 1;
        }
    } else {
        // the recurrence and the series sum are computed here, beside the decision, and committed after it: a particle the
        // decision latches keeps its state (the update of term t is dropped)
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int task = (tid - 32) + u * MID_XT;
            if (task - lane >= 2 * E4) continue;                 // warp-uniform: no task of this warp in this round
            const int side = task >= E4 ? 1 : 0;
            const int eq = task - side * E4, e = eq >> 2, q = eq & 3;
            const bool ok = task < 2 * E4 && o0 + e < N;
            double hx = 0.0;
            if (ok) {
                const double* vp = val + (side ? Gc * E4 : 0) + eq;
                const int np = side ? Gr : Gc;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four interleaved accumulators: a fixed order, short dependent chains
                int kk = 0;
                for (; kk + 4 <= np; kk += 4) {
                    a0 += vp[(size_t)kk * E4]; a1 += vp[(size_t)(kk + 1) * E4]; a2 += vp[(size_t)(kk + 2) * E4]; a3 += vp[(size_t)(kk + 3) * E4];
                }
                if (kk < np) a0 += vp[(size_t)kk * E4];
                if (kk + 1 < np) a1 += vp[(size_t)(kk + 1) * E4];
                if (kk + 2 < np) a2 += vp[(size_t)(kk + 2) * E4];
                hx = (a0 + a1) + (a2 + a3);
            }
            const double ho = __shfl_xor_sync(0xffffffffu, hx, 1);              // the other real of the complex value
            const int p = q >> 1, cmp = q & 1;
            const Cx hc = cmp ? Cx{ho, hx} : Cx{hx, ho};
            const size_t o = ((size_t)side * E + e) * NQ + 2 * p;
            if (ok) {
                const PartPass& pa = sh->spass[par].part[p];
                double2 cur = *reinterpret_cast<const double2*>(ocur + o);
                s_old[u] = cmp ? cur.y : cur.x;
                if (pa.active && !sh->sctrl.part[p].latched) {   // latched as of term t-2; the decision on t-1 is re-checked at the commit
                    double2 sum = *reinterpret_cast<const double2*>(osum + o);
                    if (pa.begin) {                              // next steady sub-step: adopt the previous sum (Taylor.f:105,:83-86);
                        flags |= 2u << (2 * u);                  // hx was computed from it (x of the chain term)
                        s_psi[u] = cmp ? sum.y : sum.x;
                        cur = sum;
                        const Cx s0 = cmul({pa.s_re, pa.s_im}, {sum.x, sum.y});
                        sum = make_double2(s0.re, s0.im);
                    }
                    Cx y = cmul({pa.alpha_re, pa.alpha_im}, hc);
                    if (pa.three_term) {
                        const Cx bc = cmul({pa.beta_re, pa.beta_im}, {cur.x, cur.y});
                        y.re += bc.re; y.im += bc.im;
                        if (pa.gamma != 0.0) {
                            const double2 prv = *reinterpret_cast<const double2*>(oprv + o);
                            y.re += pa.gamma * prv.x; y.im += pa.gamma * prv.y;
                        }
                    }
                    Cx tt = y;
                    if (pa.scale_term) tt = cmul({pa.c_re, pa.c_im}, y);
                    const double nw_re = sum.x + tt.re, nw_im = sum.y + tt.im;
                    const double dx = nw_re - sum.x, dy = nw_im - sum.y;
                    s_mag[u] = dx * dx + dy * dy;                // |new - old|^2 (isConverged, Taylor.f:290-303); root after the max
                    flags |= 1u << (2 * u);
                    s_prv[u] = cmp ? cur.y : cur.x;
                    // what the next product multiplies: the new vector, or (speculatively) the sum the next sub-step starts from
                    s_cur[u] = pa.chain ? (cmp ? nw_im : nw_re) : (cmp ? y.im : y.re);
                    s_sum[u] = cmp ? nw_im : nw_re;
                }
            }
        }
    }
    DYB_MSTAMP(11);
    __syncthreads();
    if (sh->sctrl.part[0].latched && sh->sctrl.part[1].latched) return true;
    DYB_MSTAMP(3);

    if (w > 0) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int task = (tid - 32) + u * MID_XT;
            const int side = task >= E4 ? 1 : 0;
            const int eq = task - side * E4, e = eq >> 2, q = eq & 3;
            if (task < 2 * E4 && o0 + e < N) {
                const int p = q >> 1, cmp = q & 1;
                const size_t o = ((size_t)side * E + e) * NQ + 2 * p;
                const bool upd = ((flags >> (2 * u)) & 1u) && !sh->sctrl.part[p].latched;
                if (upd) {
                    oprv[o + cmp] = s_prv[u]; ocur[o + cmp] = s_cur[u]; osum[o + cmp] = s_sum[u];
                    if ((flags >> (2 * u)) & 2u) opsi[o + cmp] = s_psi[u];
                }
                if (cmp == 0) omag[((size_t)side * E + e) * 2 + p] = upd ? s_mag[u] : 0.0;
                ll_store(P.xx + (((size_t)((t + 1) & 1) * 2 + side) * N + (o0 + e)) * NQ + q, upd ? s_cur[u] : s_old[u], ep);
            }
        }
        asm volatile("bar.arrive 1, 256;" ::: "memory");         // the owner state of term t is complete (warp 0 waits for it)
        return false;
    }

    // ---- warp 0: scalars of the owned indices
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (blockIdx.x < P.n_own) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};                      // max_b, max_k, dot_re, dot_im of particle (lane & 1)
        for (int idx = lane; idx < E * 2; idx += 32) {
            const int e = idx >> 1, p = idx & 1;
            if (o0 + e < N) {
                const double2 k = *reinterpret_cast<const double2*>(osum + (size_t)e * NQ + 2 * p);
                const double2 b = *reinterpret_cast<const double2*>(osum + ((size_t)E + e) * NQ + 2 * p);
                v[0] = fmax(v[0], omag[((size_t)E + e) * 2 + p]); v[1] = fmax(v[1], omag[(size_t)e * 2 + p]);
                v[2] += b.x * k.x + b.y * k.y;                   // conj(bra) * ket
                v[3] += b.x * k.y - b.y * k.x;
            }
        }
#pragma unroll
        for (int off = 2; off < 32; off <<= 1) {                 // lanes of equal particle (lane bit 0)
            v[0] = fmax(v[0], __shfl_xor_sync(0xffffffffu, v[0], off)); v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], off));
            v[2] += __shfl_xor_sync(0xffffffffu, v[2], off);            v[3] += __shfl_xor_sync(0xffffffffu, v[3], off);
        }
        if (lane < 2) {
            ulonglong2* dst = P.sc + ((size_t)(t & 3) * G + blockIdx.x) * 8 + lane * 4;
            ll_store(dst + 0, sqrt(v[0]), ep); ll_store(dst + 1, sqrt(v[1]), ep);
            ll_store(dst + 2, v[2], ep);       ll_store(dst + 3, v[3], ep);
        }
    }
    return false;
}

// Consumer (warps 1..7): the Cnp ket entries and the R bra entries the next product multiplies -> sx (sxk [Cnp][NQ] then
// sxb [R][NQ], contiguous: word wi of the list is double wi of that array)
template <int R>
__device__ __noinline__ void mid_consume(MidShared* sh, double* sx, int t) {
    const MidParams& P = sh->P;
    const int ct = threadIdx.x - 32;
    const int Cnp = P.Cnp, N = P.N;
    const int bi = blockIdx.x / P.Gc, bj = blockIdx.x % P.Gc;
    const int row0 = bi * R, col0 = bj * Cnp;
    const unsigned ep = P.epoch0 + t + 1;
    const ulonglong2* xk_src = P.xx + (((size_t)((t + 1) & 1) * 2 + 0) * N + col0) * NQ;
    const ulonglong2* xb_src = P.xx + (((size_t)((t + 1) & 1) * 2 + 1) * N + row0) * NQ - (size_t)Cnp * NQ;    // indexed by wi
    const int nwk = min(Cnp, max(0, N - col0)) * NQ, nwb = min(R, max(0, N - row0)) * NQ, c4 = Cnp * NQ;
    unsigned pend = 0;
    [[maybe_unused]] int ll_rounds = 0;
#pragma unroll
    for (int u = 0; u < MID_XU; ++u) {
        const int wi = ct + u * MID_XT;
        if (wi < nwk || (unsigned)(wi - c4) < (unsigned)nwb) pend |= 1u << u;
    }
#define DYB_X_ADDR(u) (((ct + (u) * MID_XT) < c4 ? xk_src : xb_src) + (ct + (u) * MID_XT))
#define DYB_X_EP(u) ep
#define DYB_X_OUT(u, v) sx[ct + (u) * MID_XT] = (v)
    DYB_LL_POLL(MID_XU, pend, DYB_X_EP, DYB_X_ADDR, DYB_X_OUT, (void)0);
#undef DYB_X_ADDR
#undef DYB_X_EP
#undef DYB_X_OUT
}

template <int WR>        // row groups of 256 rows per CTA; WC = 8 / WR column groups
__global__ void __launch_bounds__(MID_THREADS, 1)
mid_series_kernel_t(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MidParams P)
{
    constexpr int WC = MID_WARPS / WR, TC = WC * MID_CPW, R = WR * MID_SUB;
    static_assert(R * TC * 8 == MID_STAGE_BYTES, "stage size");
    extern __shared__ __align__(128) uint8_t msm[];
    const MidSmem L(P.ST, R, P.Cnp, P.E, P.Gr + P.Gc, P.n_own, P.tab16);
    uint64_t* bar_full  = reinterpret_cast<uint64_t*>(msm + L.bars);
    uint64_t* bar_empty = bar_full + 8;
    double* U    = reinterpret_cast<double*>(msm + L.U);
    double* sxk  = reinterpret_cast<double*>(msm + L.xk);
    double* sxb  = reinterpret_cast<double*>(msm + L.xb);
    double* ocur = reinterpret_cast<double*>(msm + L.own);     // [side][E][NQ]
    __shared__ MidShared sh;

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int wr = w / WC, wc = w % WC;
    const int Gr = P.Gr, Gc = P.Gc, Cnp = P.Cnp, NT = P.NT, ST = P.ST, N = P.N, E = P.E;
    [[maybe_unused]] const int G = Gr * Gc;                    // (phase stamps of the diagnostic build)
    const int bi = blockIdx.x / Gc, bj = blockIdx.x % Gc;
    const int row0 = bi * R, col0 = bj * Cnp;
    const int o0 = blockIdx.x * E;                             // first owned index
    const int total_tiles = P.n_steps * NT;
    double* oprv = ocur + 2 * E * NQ;
    double* osum = oprv + 2 * E * NQ;
    double* opsi = osum + 2 * E * NQ;                          // start vector of the sub-step in progress (handed back after a failure)

    uint64_t policy;
    if (P.l2_frac > 0.f) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(policy) : "f"(P.l2_frac));
    else policy = policy_evict_first();

    // tile q of the endless sequence (term q / NT, column tile q % NT) -> stage q % ST
    auto issue = [&](int stage, int jt) {
        mbar_arrive_expect_tx(&bar_full[stage], MID_STAGE_BYTES);
        tma_load_3d(msm + (size_t)stage * MID_STAGE_BYTES, &tmap, &bar_full[stage], 0, bi * WR, col0 + jt * TC, policy);
    };

    if (tid == 0) {
        prefetch_tensormap(&tmap);
        for (int s = 0; s < ST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], MID_WARPS); }
        fence_barrier_init();
        for (int q = 0; q < ST && q < total_tiles; ++q) issue(q, q % NT);
        sh.P = P;
        sh.sctrl = *P.ctrl;
        sh.stop_chain = 0;
    }

    // ---- consumer copy of x for the Cnp ket entries and the R bra entries this CTA multiplies
    for (int f = tid; f < (Cnp + R) * 2; f += MID_THREADS) {
        const int side = f >= Cnp * 2;
        const int ff = side ? f - Cnp * 2 : f;
        const int g = (side ? row0 : col0) + (ff >> 1);
        double2 cur = make_double2(0.0, 0.0);
        if (g < N) cur = *reinterpret_cast<const double2*>((side ? P.x0b : P.x0k) + (size_t)g * NQ + 2 * (ff & 1));
        *reinterpret_cast<double2*>((side ? sxb : sxk) + (size_t)ff * 2) = cur;
    }
    // ---- owner copy of the state of the E indices this CTA owns
    for (int f = tid; f < 2 * E * 2; f += MID_THREADS) {
        const int side = f >= E * 2;
        const int ff = side ? f - E * 2 : f;
        const int g = o0 + (ff >> 1);
        double2 cur = make_double2(0.0, 0.0), sum = cur;
        if (g < N) {
            cur = *reinterpret_cast<const double2*>((side ? P.x0b : P.x0k) + (size_t)g * NQ + 2 * (ff & 1));
            sum = *reinterpret_cast<const double2*>((side ? P.sum_b : P.sum_k) + (size_t)g * NQ + 2 * (ff & 1));
        }
        const size_t o = (size_t)side * E * NQ + (size_t)ff * 2;
        *reinterpret_cast<double2*>(ocur + o) = cur;
        *reinterpret_cast<double2*>(oprv + o) = make_double2(0.0, 0.0);
        *reinterpret_cast<double2*>(osum + o) = sum;
        *reinterpret_cast<double2*>(opsi + o) = cur;
    }

    // ---- tables of the owner's collect (word w = k * E4 + e * 4 + q; source k: ket partial of block column k < Gc, bra partial
    // of block row k - Gc; thread tid takes the words tid + 256 j), built once per launch: the per-term address arithmetic
    // was 1300 instructions per thread (3400 cycles per term).  Offsets are inside one parity of pk / pb.
    {
        const int E4 = E * NQ, W = (Gc + Gr) * E4, Wr = (W + MID_THREADS - 1) / MID_THREADS * MID_THREADS;
        int* kt = reinterpret_cast<int*>(msm + L.tab);
        if (P.tab16) {
            int* rt = kt + (Gc + Gr);
            unsigned short* wtab = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(kt) + ((Gc + Gr + 2 * E4) * 4 + 127) / 128 * 128);
            for (int k = tid; k < Gc + Gr; k += MID_THREADS) kt[k] = k < Gc ? k * R * NQ : (k - Gc) * Cnp * NQ;
            for (int f = tid; f < 2 * E4; f += MID_THREADS) {
                const int side = f >= E4, rem = f - side * E4, g = o0 + (rem >> 2), q = rem & 3;
                int v = 0;
                if (g < N) {
                    if (!side) { const int br = g / R; v = (br * Gc * R + (g - br * R)) * NQ + q; }           // pk[par][br][k][rr][q]
                    else       { const int bc = g / Cnp; v = (bc * Gr * Cnp + (g - bc * Cnp)) * NQ + q; }     // pb[par][bc][k - Gc][cc][q]
                }
                rt[f] = v;
            }
            for (int wj = tid; wj < Wr; wj += MID_THREADS) {
                int v = MID_NO_WORD;
                if (wj < W) { const int k = wj / E4, rem = wj - k * E4; if (o0 + (rem >> 2) < N) v = (k << 8) | rem; }
                wtab[wj] = (unsigned short)v;
            }
        } else {
            for (int wj = tid; wj < Wr; wj += MID_THREADS) {
                int v = MID_NO_WORD32;
                if (wj < W) {
                    const int k = wj / E4, rem = wj - k * E4, g = o0 + (rem >> 2), q = rem & 3;
                    if (g < N) {
                        if (k < Gc) { const int br = g / R; v = ((br * Gc + k) * R + (g - br * R)) * NQ + q; }
                        else        { const int bc = g / Cnp; v = ~(((bc * Gr + (k - Gc)) * Cnp + (g - bc * Cnp)) * NQ + q); }
                    }
                }
                kt[wj] = v;
            }
        }
    }

    bool decided_all = false;
    constexpr int PW = sizeof(PassParams) / 8;
    double pass_word = 0.0;
    if (tid < PW && P.n_steps > 0) pass_word = reinterpret_cast<const double*>(P.passes)[tid];
    __syncthreads();
    const bool act0[2] = {!sh.sctrl.part[0].latched, !sh.sctrl.part[1].latched};      // particles that take part in this launch

    // ring cursors (uniform over the CTA): tile being computed / retired / issued at retirement
    int qc = 0, sc = 0, phc = 0;                 // computed:  index, stage, phase
    int sr = 0, phr = 0;                         // retired:   stage, phase (index qc - 1 when used)
    int jn = ST % NT;                            // column tile of the tile issued at the next retirement (index + ST)
    auto retire = [&](int rq) {                  // all 8 warps are done with tile rq: refill its stage with tile rq + ST
        if ((rq & (MID_WARPS - 1)) == w) {
            if (lane == 0) {
                mbar_wait_or_trap(&bar_empty[sr], uint32_t(phr));
                if (rq + ST < total_tiles) issue(sr, jn);
            }
            __syncwarp();
        }
        if (++sr == ST) { sr = 0; phr ^= 1; }
        if (++jn == NT) jn = 0;
    };

    int t = 0;
    for (; t < P.n_steps; ++t) {
        if (sh.sctrl.part[0].latched && sh.sctrl.part[1].latched) { decided_all = true; break; }
        if (tid < PW) {                                          // this term's parameters were fetched one term ahead
            reinterpret_cast<double*>(&sh.spass[t & 1])[tid] = pass_word;
            if (t + 1 < P.n_steps) pass_word = reinterpret_cast<const double*>(P.passes + t + 1)[tid];
        }
        const int par = t & 1;
        const unsigned ep = P.epoch0 + t + 1;
        DYB_MSTAMP(0);

        // ---------------------------------------------------------------- 1. both products of the block, streamed
        {
            double acc[MID_MPT][2][NQ], xb[MID_MPT][2][NQ];
#pragma unroll
            for (int m = 0; m < MID_MPT; ++m) {
                const double2* xp = reinterpret_cast<const double2*>(sxb + (size_t)(wr * MID_SUB + 64 * m + 2 * lane) * NQ);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double2 v0 = xp[2 * e], v1 = xp[2 * e + 1];
                    xb[m][e][0] = v0.x; xb[m][e][1] = v0.y; xb[m][e][2] = v1.x; xb[m][e][3] = v1.y;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) acc[m][e][q] = 0.0;
                }
            }
            double pvo[MID_CPW * NQ];                            // bra partials of the previous tile, reduced beside the FMAs of this one
#pragma unroll
            for (int i = 0; i < MID_CPW * NQ; ++i) pvo[i] = 0.0;
            double* Uw = U + ((size_t)wr * Cnp + wc * MID_CPW + (lane >> 4)) * NQ + ((lane >> 2) & 3);
            for (int j = 0; j < NT; ++j) {
                mbar_wait_or_trap(&bar_full[sc], uint32_t(phc));
                const double* sH = reinterpret_cast<const double*>(msm + (size_t)sc * MID_STAGE_BYTES);
                // The five butterfly levels of tile j-1 (transpose_reduce<8>, same order) are issued BETWEEN the four FMA blocks
                // of tile j: a shuffle's latency is covered by the 32 independent FMAs that follow it in the instruction stream.
                double2 h0[MID_MPT], h1[MID_MPT];
                double xk0[NQ], xk1[NQ];
                {
                    const double2* xq = reinterpret_cast<const double2*>(sxk + (size_t)(j * TC + wc * MID_CPW) * NQ);
                    const double2 a01 = xq[0], a23 = xq[1], b01 = xq[2], b23 = xq[3];
                    xk0[0] = a01.x; xk0[1] = a01.y; xk0[2] = a23.x; xk0[3] = a23.y;
                    xk1[0] = b01.x; xk1[1] = b01.y; xk1[2] = b23.x; xk1[3] = b23.y;
                    const double2* hp = reinterpret_cast<const double2*>(sH + (size_t)((wc * MID_CPW) * WR + wr) * MID_SUB) + lane;
#pragma unroll
                    for (int m = 0; m < MID_MPT; ++m) { h0[m] = hp[m * 32]; h1[m] = hp[(size_t)WR * (MID_SUB / 2) + m * 32]; }
                }
                double pa[NQ], pa1[NQ], pb_[NQ], pb1[NQ];
                double ra[4], rb[2], rc;
                {
                    const bool hi = lane & 16;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double keep = hi ? pvo[4 + i] : pvo[i], send = hi ? pvo[i] : pvo[4 + i];
                        ra[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
                }
                mid_fma_half<0>(acc, xb, h0, xk0, pa, pa1);
                {
                    const bool hi = lane & 8;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double keep = hi ? ra[2 + i] : ra[i], send = hi ? ra[i] : ra[2 + i];
                        rb[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                }
                mid_fma_half<2>(acc, xb, h0, xk0, pa, pa1);
                {
                    const bool hi = lane & 4;
                    const double keep = hi ? rb[1] : rb[0], send = hi ? rb[0] : rb[1];
                    rc = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                mid_fma_half<0>(acc, xb, h1, xk1, pb_, pb1);
                rc += __shfl_xor_sync(0xffffffffu, rc, 2);
                mid_fma_half<2>(acc, xb, h1, xk1, pb_, pb1);
                rc += __shfl_xor_sync(0xffffffffu, rc, 1);                        // lane holds value (lane >> 2) of tile j-1
                if ((lane & 3) == 0 && j > 0) Uw[(size_t)(j - 1) * TC * NQ] = rc;
#pragma unroll
                for (int q = 0; q < NQ; ++q) { pvo[q] = pa[q] + pa1[q]; pvo[NQ + q] = pb_[q] + pb1[q]; }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[sc]);      // release: stage reads are done
                if (j > 0) retire(qc - 1);
                ++qc;
                if (++sc == ST) { sc = 0; phc ^= 1; }
            }
            {
                const double tot = transpose_reduce<MID_CPW * NQ>(pvo, lane);
                if ((lane & 3) == 0) Uw[(size_t)(NT - 1) * TC * NQ] = tot;
            }
            retire(qc - 1);
            DYB_MSTAMP(1);
            __syncthreads();                                     // bra partials of all row groups are in U

            // bra partial of block column bj from block row bi -> pb[par][bj][bi][.]
            {
                ulonglong2* dst = P.pb + (((size_t)par * Gc + bj) * Gr + bi) * Cnp * NQ;
                for (int idx = tid; idx < Cnp * NQ; idx += MID_THREADS) {
                    double v = U[idx];
#pragma unroll
                    for (int r2 = 1; r2 < WR; ++r2) v += U[(size_t)r2 * Cnp * NQ + idx];
                    ll_store(dst + idx, v, ep);
                }
            }
            __syncthreads();                                     // U is free for the ket reduction

            // ket sums of the WC column groups: pairwise tree through shared memory, fixed order
            double2* kred = reinterpret_cast<double2*>(U);
#pragma unroll
            for (int s = 1, r = 0; s < WC; s <<= 1, ++r) {
                double2* slot = kred + ((size_t)(wr * (WC >> 1) + (wc >> (r + 1))) * MID_SUB * 2);
                if ((wc & (2 * s - 1)) == s) {
#pragma unroll
                    for (int m = 0; m < MID_MPT; ++m)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            slot[((m * 2 + e) * 2 + 0) * 32 + lane] = make_double2(acc[m][e][0], acc[m][e][1]);
                            slot[((m * 2 + e) * 2 + 1) * 32 + lane] = make_double2(acc[m][e][2], acc[m][e][3]);
                        }
                }
                __syncthreads();
                if ((wc & (2 * s - 1)) == 0) {
#pragma unroll
                    for (int m = 0; m < MID_MPT; ++m)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double2 a = slot[((m * 2 + e) * 2 + 0) * 32 + lane], b = slot[((m * 2 + e) * 2 + 1) * 32 + lane];
                            acc[m][e][0] += a.x; acc[m][e][1] += a.y; acc[m][e][2] += b.x; acc[m][e][3] += b.y;
                        }
                }
                if (2 * s < WC) __syncthreads();
            }
            // ket partial of block row bi from block column bj -> pk[par][bi][bj][.]: rows through shared memory (the region of
            // U a wc == 0 warp writes is its own or already consumed), then 512 contiguous bytes per store instruction
            if (wc == 0) {
                __syncwarp();                                    // the lanes of this warp have read their part of the tree slot it overwrites
                double2* st = reinterpret_cast<double2*>(U) + (size_t)(wr * MID_SUB + 2 * lane) * 2;
#pragma unroll
                for (int m = 0; m < MID_MPT; ++m)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        st[(m * 64 + e) * 2 + 0] = make_double2(acc[m][e][0], acc[m][e][1]);
                        st[(m * 64 + e) * 2 + 1] = make_double2(acc[m][e][2], acc[m][e][3]);
                    }
            }
            __syncthreads();
            {
                ulonglong2* dst = P.pk + ((((size_t)par * Gr + bi) * Gc + bj) * R) * NQ;
#pragma unroll
                for (int i = 0; i < R * NQ / MID_THREADS; ++i) ll_store(dst + i * MID_THREADS + tid, U[i * MID_THREADS + tid], ep);
            }
        }
        DYB_MSTAMP(2);

        // (the ket rows staged in U are overwritten by what the owner collects: mid_owner synchronises after its first request)

        // ---------------------------------------------------------------- 2. owner: collect, decide on term t-1, update, publish
        if (mid_owner<R>(&sh, U, ocur, t)) { decided_all = true; break; }
        DYB_MSTAMP(4);

        // ---------------------------------------------------------------- 3. consumer: the entries the next product multiplies
        if (w > 0 && t + 1 < P.n_steps) mid_consume<R>(&sh, sxk, t);
        DYB_MSTAMP(5);
        __syncthreads();
        DYB_MSTAMP(6);
    }

    // ---- decision on the last term, unless the series was decided on the way
    if (!decided_all && t > 0) mid_decide(&sh, t - 1, true);

    // ---- never exit with bulk copies in flight to our shared memory: tiles qc .. min(total, qc + ST) - 1 were issued
    if (tid == 0) {
        const int issued = min(total_tiles, qc + ST);
        for (int q = qc; q < issued; ++q) mbar_wait_or_trap(&bar_full[q % ST], uint32_t((q / ST) & 1));
    }

    // ---- results of the owned indices
    // a particle that failed a steady sub-step hands back the start vector of that sub-step (= the last accepted sum)
    for (int f = tid; f < 2 * E * 2; f += MID_THREADS) {
        const int side = f >= E * 2;
        const int ff = side ? f - E * 2 : f;
        const int e = ff >> 1, p = ff & 1, g = o0 + e;
        if (g >= N || !act0[p]) continue;
        const bool failed = sh.sctrl.part[p].latched && !sh.sctrl.part[p].ok;
        const size_t o = ((size_t)side * E + e) * NQ + 2 * p;
        const double2 v = *reinterpret_cast<const double2*>((failed ? opsi : osum) + o);
        *reinterpret_cast<double2*>((side ? P.sum_b : P.sum_k) + (size_t)g * NQ + 2 * p) = v;
    }
    if (blockIdx.x == 0 && tid == 0) {
        sh.sctrl.all_latched = (sh.sctrl.part[0].latched && sh.sctrl.part[1].latched) ? 1 : 0;
        sh.sctrl.block_counter = 0u;
        *P.ctrl = sh.sctrl;
    }
}

}  // namespace dyb
