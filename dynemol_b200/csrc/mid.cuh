// mid.cuh -- mid-size operators (roughly 1800 < N < 6000): ONE cooperative launch per series, H' STREAMED per term.
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
//     columns (32 KiB) and RUNS ACROSS THE TERMS: H' does not change, so the first tiles of term t+1 are in flight while the
//     grid barrier and the gather of term t run; the L2 policy keeps as much of H' as fits resident (evict_last on a
//     fraction of the lines, evict_first on the rest), so that most of a pass is served by the L2, not by HBM;
//   * the tile engine of the dual product (matvec.cuh): 8 rows x 2 columns per thread and tile, ket sums in registers,
//     bra sums by a lane butterfly (transpose_reduce) -- one pass serves H'x_ket and H'^T x_bra of electron and hole;
//   * ONE grid barrier per term and the redundant-gather protocol of resident.cuh: after the barrier every CTA rebuilds,
//     in the same order (bit-identical copies), exactly the vector entries it multiplies next -- x_ket on its columns from
//     the Gc ket partials of that block row, x_bra on its rows from the Gr bra partials of that block column -- and applies
//     the recurrence / series sum to its copy of the state (x, x_prev, sum: shared memory);
//   * the convergence scalars come from the CTAs whose row and column ranges intersect (they hold bra AND ket sums of
//     those indices) and are consumed one term late, with the decision code of the other two paths (decide_particle).
//
// Same PassParams / Ctrl contract as resident.cuh: propagate_series does not know which kernel ran.  Chained steady
// sub-steps (PartPass::begin / chain) are supported; the reference-GPU term test (PartPass::test_gpu) is not -- those parity
// modes stay on the two-launch path.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"
#include "matvec.cuh"
#include "resident.cuh"

namespace dyb {

constexpr int MID_THREADS     = 256;
constexpr int MID_WARPS       = 8;
constexpr int MID_SUB         = 256;        // rows per warp (8 per thread: lane l owns rows 64m + 2l + {0,1}, m = 0..3)
constexpr int MID_MPT         = 4;
constexpr int MID_CPW         = 2;          // columns per warp and tile
constexpr int MID_STAGE_BYTES = 32768;      // R x TC x 8 B with R*TC = 4096 for every WR
constexpr int MID_MAX_ST      = 5;
constexpr int MID_U_BYTES     = 32768;      // union region: bra partials of the row groups / ket tree reduction / term magnitudes
constexpr int MID_MAX_DIAG    = 160;
constexpr int MID_GB          = 4;          // gather: slots a thread keeps in flight
constexpr int MID_SMEM_MAX    = 227 * 1024 - 2048;

struct MidParams {
    int N, Gr, Gc, Cnp, NT, ST;             // grid, block width (multiple of TC), tiles per term, ring depth
    int lslk, lslb;                         // log2 of the lanes that share a gather task (ket: Gc partials, bra: Gr partials)
    int nd;                                 // CTAs whose row and column ranges intersect, in blockIdx order
    float l2_frac;                          // fraction of the H' lines loaded with L2::evict_last (0: plain evict_first stream)
    const double* x0k; const double* x0b;   // starting vectors (quads), written by series_init_kernel
    double* sum_b; double* sum_k;           // in: series sums at the start; out: at the latch / end of the series
    double* pk; double* pb;                 // [2][Gr][Gc][R][NQ] / [2][Gc][Gr][Cnp][NQ] partial products (parity of the term first)
    double* dscal;                          // [2][grid][8] scalars of the intersecting CTAs
    double* psi_store;                      // [2][N][NQ] start vector of the sub-step in progress (ket, bra), needed after a failure
    Ctrl* ctrl;
    const PassParams* passes; int n_steps;
    unsigned long long* gbar;
    long long* prof;                        // DYB_SERIES_PROF builds: [32][grid][8] clock64 stamps (else null)
    int diag[MID_MAX_DIAG];
};

#ifdef DYB_SERIES_PROF
#define DYB_MSTAMP(i) do { if (threadIdx.x == 0 && t < 32) P.prof[((size_t)t * G + blockIdx.x) * 8 + (i)] = clock64(); } while (0)
#else
#define DYB_MSTAMP(i) do { } while (0)
#endif

struct MidSmem {                            // dynamic shared memory carve-up (byte offsets)
    int bars, U, xk, prvk, sumk, xb, prvb, sumb, total;
    __host__ __device__ MidSmem(int ST, int R, int Cnp) {
        bars = ST * MID_STAGE_BYTES;
        U    = bars + 128;
        xk   = U + MID_U_BYTES;
        prvk = xk + Cnp * 32;  sumk = prvk + Cnp * 32;
        xb   = sumk + Cnp * 32;
        prvb = xb + R * 32;    sumb = prvb + R * 32;
        total = sumb + R * 32;
    }
};

// 64 FMAs of one column: the thread's 8 rows against the 4 reals of x_ket (-> acc) and of x_bra (-> p)
__device__ __forceinline__ void mid_fma_column(double (&acc)[MID_MPT][2][NQ], const double (&xb)[MID_MPT][2][NQ],
                                               const double2 (&h)[MID_MPT], const double (&xk)[NQ], double* p) {
    double p1[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) { p[q] = h[0].x * xb[0][0][q]; p1[q] = h[0].y * xb[0][1][q]; }
#pragma unroll
    for (int m = 0; m < MID_MPT; ++m) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            acc[m][0][q] = fma(h[m].x, xk[q], acc[m][0][q]);
            acc[m][1][q] = fma(h[m].y, xk[q], acc[m][1][q]);
            if (m > 0) { p[q] = fma(h[m].x, xb[m][0][q], p[q]); p1[q] = fma(h[m].y, xb[m][1][q], p1[q]); }
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) p[q] += p1[q];
}

template <int WR>        // row groups of 256 rows per CTA; WC = 8 / WR column groups
__global__ void __launch_bounds__(MID_THREADS, 1)
mid_series_kernel_t(const __grid_constant__ CUtensorMap tmap, const MidParams P)
{
    constexpr int WC = MID_WARPS / WR, TC = WC * MID_CPW, R = WR * MID_SUB;
    static_assert(R * TC * 8 == MID_STAGE_BYTES, "stage size");
    extern __shared__ __align__(128) uint8_t msm[];
    const MidSmem L(P.ST, R, P.Cnp);
    uint64_t* bar_full  = reinterpret_cast<uint64_t*>(msm + L.bars);
    uint64_t* bar_empty = bar_full + 8;
    double* U    = reinterpret_cast<double*>(msm + L.U);
    double* sxk  = reinterpret_cast<double*>(msm + L.xk);
    double* sxb  = reinterpret_cast<double*>(msm + L.xb);
    __shared__ Ctrl       sctrl;
    __shared__ PassParams spass[2];
    __shared__ double     fin[8];
    __shared__ double     wred[MID_WARPS][8];
    __shared__ int        stop_chain;

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int wr = w / WC, wc = w % WC;
    const int Gr = P.Gr, Gc = P.Gc, Cnp = P.Cnp, NT = P.NT, ST = P.ST, N = P.N;
    const int G = Gr * Gc;
    const int bi = blockIdx.x / Gc, bj = blockIdx.x % Gc;
    const int row0 = bi * R, col0 = bj * Cnp;
    const int total_tiles = P.n_steps * NT;
    // indices whose bra AND ket entries this CTA holds
    const int i0 = max(row0, col0), i1 = min(min(row0 + R, col0 + Cnp), N);
    const bool diag = i0 < i1;

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
        sctrl = *P.ctrl;
        stop_chain = 0;
    }

    // ---- this CTA's copy of the state: x (= cur), x_prev, sum for its Cnp ket entries and its R bra entries
    for (int f = tid; f < (Cnp + R) * 2; f += MID_THREADS) {
        const int side = f >= Cnp * 2;
        const int ff = side ? f - Cnp * 2 : f;
        const int e = ff >> 1, p = ff & 1;
        const int g = (side ? row0 : col0) + e;
        double2 cur = make_double2(0.0, 0.0), sum = cur;
        if (g < N) {
            cur = *reinterpret_cast<const double2*>((side ? P.x0b : P.x0k) + (size_t)g * NQ + 2 * p);
            sum = *reinterpret_cast<const double2*>((side ? P.sum_b : P.sum_k) + (size_t)g * NQ + 2 * p);
            if (g >= i0 && g < i1) __stcg(reinterpret_cast<double2*>(P.psi_store + ((size_t)side * N + g) * NQ + 2 * p), cur);
        }
        double* st = reinterpret_cast<double*>(msm + (side ? L.xb : L.xk)) + (size_t)e * NQ + 2 * p;
        const int n_e = side ? R : Cnp;
        *reinterpret_cast<double2*>(st) = cur;
        *reinterpret_cast<double2*>(st + (size_t)n_e * NQ) = make_double2(0.0, 0.0);
        *reinterpret_cast<double2*>(st + (size_t)2 * n_e * NQ) = sum;
    }

    unsigned long long bar_target = 0;
    bool decided_all = false;
    constexpr int PW = sizeof(PassParams) / 8;
    double pass_word = 0.0;
    if (tid < PW && P.n_steps > 0) pass_word = reinterpret_cast<const double*>(P.passes)[tid];
    __syncthreads();
    const bool act0[2] = {!sctrl.part[0].latched, !sctrl.part[1].latched};      // particles that take part in this launch

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
        if (sctrl.part[0].latched && sctrl.part[1].latched) { decided_all = true; break; }
        if (tid < PW) {                                          // this term's parameters were fetched one term ahead
            reinterpret_cast<double*>(&spass[t & 1])[tid] = pass_word;
            if (t + 1 < P.n_steps) pass_word = reinterpret_cast<const double*>(P.passes + t + 1)[tid];
        }
        const int par = t & 1;
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
            for (int j = 0; j < NT; ++j) {
                mbar_wait_or_trap(&bar_full[sc], uint32_t(phc));
                const double* sH = reinterpret_cast<const double*>(msm + (size_t)sc * MID_STAGE_BYTES);
                double pv[MID_CPW * NQ];
#pragma unroll
                for (int cc = 0; cc < MID_CPW; ++cc) {
                    const int c = wc * MID_CPW + cc;
                    const double2* xq = reinterpret_cast<const double2*>(sxk + (size_t)(j * TC + c) * NQ);
                    const double2 x01 = xq[0], x23 = xq[1];
                    const double xk[NQ] = {x01.x, x01.y, x23.x, x23.y};
                    const double2* hp = reinterpret_cast<const double2*>(sH + (size_t)(c * WR + wr) * MID_SUB) + lane;
                    double2 h[MID_MPT];
#pragma unroll
                    for (int m = 0; m < MID_MPT; ++m) h[m] = hp[m * 32];
                    mid_fma_column(acc, xb, h, xk, pv + cc * NQ);
                }
                const double tot = transpose_reduce<MID_CPW * NQ>(pv, lane);       // lane holds value (lane >> 2)
                if ((lane & 3) == 0) {
                    const int v = lane >> 2;
                    U[((size_t)wr * Cnp + j * TC + wc * MID_CPW + (v >> 2)) * NQ + (v & 3)] = tot;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[sc]);      // release: stage reads are done
                if (j > 0) retire(qc - 1);
                ++qc;
                if (++sc == ST) { sc = 0; phc ^= 1; }
            }
            retire(qc - 1);
            DYB_MSTAMP(1);
            __syncthreads();                                     // bra partials of all row groups are in U

            // bra partial of block column bj from block row bi -> pb[par][bj][bi][.]
            {
                double* dst = P.pb + (((size_t)par * Gc + bj) * Gr + bi) * Cnp * NQ;
                for (int idx = tid; idx < Cnp * 2; idx += MID_THREADS) {
                    double2 v = *reinterpret_cast<const double2*>(U + (size_t)idx * 2);
#pragma unroll
                    for (int r2 = 1; r2 < WR; ++r2) {
                        const double2 o = *reinterpret_cast<const double2*>(U + ((size_t)r2 * Cnp * NQ) + (size_t)idx * 2);
                        v.x += o.x; v.y += o.y;
                    }
                    __stcg(reinterpret_cast<double2*>(dst + (size_t)idx * 2), v);
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
            // ket partial of block row bi from block column bj -> pk[par][bi][bj][.]
            if (wc == 0) {
                double2* dst = reinterpret_cast<double2*>(P.pk + ((((size_t)par * Gr + bi) * Gc + bj) * R + wr * MID_SUB + 2 * lane) * NQ);
#pragma unroll
                for (int m = 0; m < MID_MPT; ++m)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        __stcg(dst + m * (64 * NQ / 2) + e * 2,     make_double2(acc[m][e][0], acc[m][e][1]));
                        __stcg(dst + m * (64 * NQ / 2) + e * 2 + 1, make_double2(acc[m][e][2], acc[m][e][3]));
                    }
            }
        }

        // ---------------------------------------------------------------- 2. the one grid barrier of the term
        DYB_MSTAMP(2);
        bar_target += G;
        res_grid_barrier(P.gbar, bar_target);
        DYB_MSTAMP(3);

        // ---------------------------------------------------------------- 3. decision on term t-1 (identical in every CTA)
        if (t > 0) {
            if (w == 0) {
                double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                for (int d = lane; d < P.nd; d += 32) {
                    const double2* ds = reinterpret_cast<const double2*>(P.dscal + ((size_t)((t + 1) & 1) * G + P.diag[d]) * 8);
                    const double2 a0 = __ldcg(ds), a1 = __ldcg(ds + 1), a2 = __ldcg(ds + 2), a3 = __ldcg(ds + 3);
                    v[0] = fmax(v[0], a0.x); v[1] = fmax(v[1], a0.y); v[2] += a1.x; v[3] += a1.y;
                    v[4] = fmax(v[4], a2.x); v[5] = fmax(v[5], a2.y); v[6] += a3.x; v[7] += a3.y;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double o = __shfl_xor_sync(0xffffffffu, v[q], off);
                        v[q] = ((q & 3) < 2) ? fmax(v[q], o) : v[q] + o;
                    }
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) fin[q] = v[q];
                }
            }
            __syncthreads();
            if (tid == 0 || tid == 32) {
                const int p = tid >> 5;                          // stop_chain as of the previous term: the same in every CTA
                decide_particle(sctrl.part[p], spass[(t - 1) & 1].part[p], fin + 4 * p, !stop_chain);
            }
            __syncthreads();
            if (sctrl.part[0].latched && sctrl.part[1].latched) { decided_all = true; break; }
            if (tid == 0 && ((sctrl.part[0].latched && !sctrl.part[0].ok) || (sctrl.part[1].latched && !sctrl.part[1].ok))) stop_chain = 1;
        }

        DYB_MSTAMP(4);
        // ---------------------------------------------------------------- 4. gather + recurrence + series sum
        // A gather task = (entry, particle) of one side; 2^lsl lanes split its partials (<= 8 each, fixed order), a lane
        // butterfly adds them up, the first lane applies the update to this CTA's copy of the state.
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int n_e   = side ? R : Cnp;
            const int gbase = side ? row0 : col0;
            const int lsl   = side ? P.lslb : P.lslk;
            const int SL    = 1 << lsl;
            const int n_par = side ? Gr : Gc;                    // partials per entry
            const int nslots = (n_e * 2) << lsl;
            double* xs = side ? sxb : sxk;
            double* ps = xs + (size_t)n_e * NQ;
            double* ss = ps + (size_t)n_e * NQ;
            double* mg = U + (side ? Cnp * 2 : 0);               // |new - old|^2 per (entry, particle)
#pragma unroll 1
            for (int f0 = 0; f0 < nslots; f0 += MID_THREADS * MID_GB) {
                double2 v[MID_GB][8];
                bool okv[MID_GB];
#pragma unroll
                for (int u = 0; u < MID_GB; ++u) {
                    const int f = f0 + u * MID_THREADS + tid;
                    const int sl = f & (SL - 1), p = (f >> lsl) & 1, e = f >> (lsl + 1);
                    const int g = gbase + e;
                    okv[u] = (f < nslots) && (g < N);
                    const double* src;
                    size_t stride;
                    if (side == 0) {                             // ket entry g: block row g / R, partials of the Gc block columns
                        const int br = g / R, rr = g - br * R;
                        src = P.pk + ((((size_t)par * Gr + br) * Gc) * R + rr) * NQ + 2 * p;
                        stride = (size_t)R * NQ;
                    } else {                                     // bra entry g: block column g / Cnp, partials of the Gr block rows
                        const int bc = g / Cnp, cc = g - bc * Cnp;
                        src = P.pb + ((((size_t)par * Gc + bc) * Gr) * Cnp + cc) * NQ + 2 * p;
                        stride = (size_t)Cnp * NQ;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = sl + (i << lsl);
                        v[u][i] = (okv[u] && k < n_par) ? __ldcg(reinterpret_cast<const double2*>(src + (size_t)k * stride)) : make_double2(0.0, 0.0);
                    }
                }
#pragma unroll
                for (int u = 0; u < MID_GB; ++u) {
                    double2 hx = v[u][0];
#pragma unroll
                    for (int i = 1; i < 8; ++i) { hx.x += v[u][i].x; hx.y += v[u][i].y; }
                    for (int off = 1; off < SL; off <<= 1) {
                        hx.x += __shfl_xor_sync(0xffffffffu, hx.x, off); hx.y += __shfl_xor_sync(0xffffffffu, hx.y, off);
                    }
                    const int f = f0 + u * MID_THREADS + tid;
                    const int sl = f & (SL - 1), p = (f >> lsl) & 1, e = f >> (lsl + 1);
                    if (okv[u] && sl == 0) {
                        const PartPass& pa = spass[par].part[p];
                        double mag = 0.0;
                        if (pa.active && !sctrl.part[p].latched) {
                            const size_t o = (size_t)e * NQ + 2 * p;
                            double2 cur = *reinterpret_cast<const double2*>(xs + o);
                            double2 sum = *reinterpret_cast<const double2*>(ss + o);
                            if (pa.begin) {                      // next steady sub-step: adopt the previous sum (Taylor.f:105,:83-86);
                                const int g = gbase + e;         // hx was computed from it (x of the chain term)
                                if (g >= i0 && g < i1) __stcg(reinterpret_cast<double2*>(P.psi_store + ((size_t)side * N + g) * NQ + 2 * p), sum);
                                cur = sum;
                                const Cx s0 = cmul({pa.s_re, pa.s_im}, {sum.x, sum.y});
                                sum = make_double2(s0.re, s0.im);
                            }
                            Cx y = cmul({pa.alpha_re, pa.alpha_im}, {hx.x, hx.y});
                            if (pa.three_term) {
                                const Cx bc = cmul({pa.beta_re, pa.beta_im}, {cur.x, cur.y});
                                y.re += bc.re; y.im += bc.im;
                                if (pa.gamma != 0.0) {
                                    const double2 prv = *reinterpret_cast<const double2*>(ps + o);
                                    y.re += pa.gamma * prv.x; y.im += pa.gamma * prv.y;
                                }
                            }
                            Cx tt = y;
                            if (pa.scale_term) tt = cmul({pa.c_re, pa.c_im}, y);
                            const double nw_re = sum.x + tt.re, nw_im = sum.y + tt.im;
                            const double dx = nw_re - sum.x, dy = nw_im - sum.y;
                            mag = dx * dx + dy * dy;             // |new - old|^2 (isConverged, Taylor.f:290-303); root after the max
                            *reinterpret_cast<double2*>(ps + o) = cur;
                            // what the next product multiplies: the new vector, or (speculatively) the sum the next sub-step starts from
                            *reinterpret_cast<double2*>(xs + o) = pa.chain ? make_double2(nw_re, nw_im) : make_double2(y.re, y.im);
                            *reinterpret_cast<double2*>(ss + o) = make_double2(nw_re, nw_im);
                        }
                        mg[e * 2 + p] = mag;
                    }
                }
            }
        }
        __syncthreads();
        DYB_MSTAMP(5);

        // ---------------------------------------------------------------- 5. scalars of the indices this CTA holds on both sides
        if (diag) {
            double v[4] = {0.0, 0.0, 0.0, 0.0};                  // max_b, max_k, dot_re, dot_im of particle (tid & 1)
            for (int idx = tid; idx < (i1 - i0) * 2; idx += MID_THREADS) {
                const int g = i0 + (idx >> 1), p = idx & 1;
                const int ek = g - col0, eb = g - row0;
                const double2 k = *reinterpret_cast<const double2*>(sxk + (size_t)2 * Cnp * NQ + (size_t)ek * NQ + 2 * p);
                const double2 b = *reinterpret_cast<const double2*>(sxb + (size_t)2 * R * NQ + (size_t)eb * NQ + 2 * p);
                v[0] = fmax(v[0], U[Cnp * 2 + eb * 2 + p]); v[1] = fmax(v[1], U[ek * 2 + p]);
                v[2] += b.x * k.x + b.y * k.y;                   // conj(bra) * ket
                v[3] += b.x * k.y - b.y * k.x;
            }
#pragma unroll
            for (int off = 2; off < 32; off <<= 1) {             // lanes of equal particle (lane bit 0)
                v[0] = fmax(v[0], __shfl_xor_sync(0xffffffffu, v[0], off)); v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], off));
                v[2] += __shfl_xor_sync(0xffffffffu, v[2], off);            v[3] += __shfl_xor_sync(0xffffffffu, v[3], off);
            }
            if (lane < 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) wred[w][lane * 4 + q] = v[q];
            }
            __syncthreads();
            if (tid < 8) {
                double f = wred[0][tid];
#pragma unroll
                for (int w2 = 1; w2 < MID_WARPS; ++w2) f = ((tid & 3) < 2) ? fmax(f, wred[w2][tid]) : f + wred[w2][tid];
                if ((tid & 3) < 2) f = sqrt(f);
                __stcg(P.dscal + ((size_t)par * G + blockIdx.x) * 8 + tid, f);
            }
            __syncthreads();                                     // the magnitudes in U have been read: the next product may reuse U
        }
        DYB_MSTAMP(6);
    }

    // ---- decision on the last term (one more barrier), unless the series was decided on the way
    if (!decided_all && t > 0) {
        bar_target += G;
        res_grid_barrier(P.gbar, bar_target);
        if (w == 0) {
            double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int d = lane; d < P.nd; d += 32) {
                const double2* ds = reinterpret_cast<const double2*>(P.dscal + ((size_t)((t - 1) & 1) * G + P.diag[d]) * 8);
                const double2 a0 = __ldcg(ds), a1 = __ldcg(ds + 1), a2 = __ldcg(ds + 2), a3 = __ldcg(ds + 3);
                v[0] = fmax(v[0], a0.x); v[1] = fmax(v[1], a0.y); v[2] += a1.x; v[3] += a1.y;
                v[4] = fmax(v[4], a2.x); v[5] = fmax(v[5], a2.y); v[6] += a3.x; v[7] += a3.y;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double o = __shfl_xor_sync(0xffffffffu, v[q], off);
                    v[q] = ((q & 3) < 2) ? fmax(v[q], o) : v[q] + o;
                }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q) fin[q] = v[q];
            }
        }
        __syncthreads();
        if (tid == 0 || tid == 32) { const int p = tid >> 5; decide_particle(sctrl.part[p], spass[(t - 1) & 1].part[p], fin + 4 * p); }
        __syncthreads();
    }

    // ---- never exit with bulk copies in flight to our shared memory: tiles qc .. min(total, qc + ST) - 1 were issued
    if (tid == 0) {
        const int issued = min(total_tiles, qc + ST);
        for (int q = qc; q < issued; ++q) mbar_wait_or_trap(&bar_full[q % ST], uint32_t((q / ST) & 1));
    }

    // ---- results: every index has exactly one CTA that holds both of its sums
    // a particle that failed a steady sub-step hands back the start vector of that sub-step (= the last accepted sum)
    if (diag) {
        for (int idx = tid; idx < (i1 - i0) * 2; idx += MID_THREADS) {
            const int g = i0 + (idx >> 1), p = idx & 1;
            if (!act0[p]) continue;
            const bool failed = sctrl.part[p].latched && !sctrl.part[p].ok;
            const size_t og = (size_t)g * NQ + 2 * p;
            const double2 k = failed ? __ldcg(reinterpret_cast<const double2*>(P.psi_store + og))
                                     : *reinterpret_cast<const double2*>(sxk + (size_t)2 * Cnp * NQ + (size_t)(g - col0) * NQ + 2 * p);
            const double2 b = failed ? __ldcg(reinterpret_cast<const double2*>(P.psi_store + (size_t)N * NQ + og))
                                     : *reinterpret_cast<const double2*>(sxb + (size_t)2 * R * NQ + (size_t)(g - row0) * NQ + 2 * p);
            *reinterpret_cast<double2*>(P.sum_k + og) = k;
            *reinterpret_cast<double2*>(P.sum_b + og) = b;
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        sctrl.all_latched = (sctrl.part[0].latched && sctrl.part[1].latched) ? 1 : 0;
        sctrl.block_counter = 0u;
        *P.ctrl = sctrl;
    }
}

}  // namespace dyb
