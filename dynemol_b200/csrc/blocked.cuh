// blocked.cuh -- mid-size operators (1824 < N <= 6144): the 2-D block decomposition and the one-barrier-per-term
// protocol of resident.cuh, with the block of H' STREAMED through shared memory every term instead of living there.
//
// Why: between the sizes where H' fits the shared memories (resident.cuh) and the sizes where a pass over H' takes
// hundreds of microseconds (matvec.cuh), the launch-per-term path pays ~17 us per term for its epilogue launch, the
// tall 2048-row panels (74 ket partials per row at N=4096) and the pipeline refill.  Here CTA (i, j) of a 12 x 12
// grid owns block (i, j) of H' (Bs x Bs, Bs = ceil(N/12) <= 512) and, per term,
//
//   1. streams the block in chunks of Cc columns (cp.async, 8-byte granularity so the shared-memory column stride can
//      stay odd = conflict-free from both sides; 3-stage ring that keeps running across the terms: H' does not
//      change, so the first chunks of term t+1 are in flight during the barrier and the epilogue of term t);
//      threads 0..319 accumulate the ket partial (a thread owns two rows, accumulators live across the chunks),
//      threads 320..639 produce the bra partial of the chunk's columns (row range split over thread groups, reduced
//      through shared memory behind a named barrier of the bra half only) and store it straight to L2;
//   2. ONE grid barrier;
//   3. gathers, redundantly and in the same order in every CTA that needs them, the new x_ket entries of block j and
//      x_bra entries of block i (12 partials each), applies the recurrence and the series sum.  The vectors of the
//      current term stay in shared memory; the previous term and the series sums of this CTA's entries live in a
//      CTA-private slice of global memory (L2), because 2 x 512 entries x 2 particles no longer fit the registers.
//
// Decisions, PassParams, Ctrl: exactly as in resident.cuh (diagonal CTAs publish the scalars, consumed one term late).
#pragma once
#include "common.cuh"
#include "epilogue.cuh"
#include "resident.cuh"

namespace dyb {

constexpr int BLK_THREADS  = 640;
constexpr int BLK_HALF     = 320;
constexpr int BLK_STAGES   = 3;
constexpr int BLK_MAX_BS   = 512;
constexpr int BLK_MAX_TASK = (2 * BLK_MAX_BS + BLK_HALF - 1) / BLK_HALF;    // (entry, particle) tasks per thread: 4
constexpr int BLK_SMEM_MAX = 227 * 1024 - 2048;

struct BlockedParams {
    const double* H; long long ld;        // column-major H' (N x N)
    int N, Gd, Bs, ldS, Cc, n_chunks;     // grid side, block size, odd smem column stride, chunk columns, chunks per block
    const double* x0k; const double* x0b; // starting vectors (quads), written by series_init_kernel
    double* sum_b; double* sum_k;         // in: series sums at the start; out: at the latch / end of the series
    double* pk; double* pb;               // [2][Gd][Gd][Bs][NQ] partial products (parity of the term first)
    double* dscal;                        // [2][Gd][8] scalars of the diagonal CTAs
    double* st_prv; double* st_sum;       // CTA-private state [G][2][Bs][NQ]: previous vector, series sum
    double* st_mag;                       // CTA-private [G][2][Bs][2]: |new - old|^2 of the last term
    Ctrl* ctrl;
    const PassParams* passes; int n_steps;
    unsigned long long* gbar;
};

struct BlockedSmem {                      // dynamic shared memory carve-up (offsets in doubles)
    int stage_len, xk, xb, partk, partb, total, NGk, NGb;
    __host__ __device__ BlockedSmem(int Bs, int ldS, int Cc) {
        stage_len = (Cc * ldS + 1) & ~1;
        xk = BLK_STAGES * stage_len;
        xb = xk + Bs * NQ;
        partk = xb + Bs * NQ;
        const int Bh = (Bs + 1) >> 1, Bo = (Bh + 31) & ~31;
        NGk = BLK_HALF / Bo; if (NGk < 1) NGk = 1;
        NGb = BLK_HALF / Cc;
        partb = partk + (NGk > 1 ? NGk * Bs * NQ : 0);
        total = partb + NGb * Cc * NQ;
    }
    __host__ __device__ size_t bytes() const { return (size_t)total * 8; }
};

__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned n = valid ? 8u : 0u;                          // src-size 0: the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory"); }
__device__ __forceinline__ void bar_sync_bra_half() { asm volatile("bar.sync 1, %0;" ::"n"(BLK_HALF) : "memory"); }

__global__ void __launch_bounds__(BLK_THREADS, 1)
blocked_series_kernel(const BlockedParams R)
{
    extern __shared__ __align__(16) double bsm[];
    const BlockedSmem L(R.Bs, R.ldS, R.Cc);
    __shared__ Ctrl       sctrl;
    __shared__ PassParams spass[2];
    __shared__ double     fin[8];
    __shared__ double     wred[BLK_HALF / 32][8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Gd = R.Gd, Bs = R.Bs, ldS = R.ldS, N = R.N, Cc = R.Cc, n_chunks = R.n_chunks;
    const int bi = blockIdx.x / Gd, bj = blockIdx.x % Gd;
    const int G = Gd * Gd;
    const bool diag = (bi == bj);
    const int side = tid / BLK_HALF, tl = tid - side * BLK_HALF;     // 0: ket half, 1: bra half
    double* const xsd = bsm + (side ? L.xb : L.xk);
    double* const partk = bsm + L.partk;
    double* const partb = bsm + L.partb;
    const int NGk = L.NGk, NGb = L.NGb;

    // ---- CTA-private state and the (entry, particle) tasks of this thread: task = tl + 320*m -> entry task>>1, particle task&1
    const int tblock = side ? bi : bj;
    const int n_task = 2 * Bs;
    const int tp = tl & 1;                                       // BLK_HALF is even: the particle of all this thread's tasks
    double* const my_prv = R.st_prv + ((size_t)(blockIdx.x * 2 + side) * Bs) * NQ;
    double* const my_sum = R.st_sum + ((size_t)(blockIdx.x * 2 + side) * Bs) * NQ;
    double* const my_mag = R.st_mag + ((size_t)(blockIdx.x * 2 + side) * Bs) * 2;
#pragma unroll
    for (int m = 0; m < BLK_MAX_TASK; ++m) {
        const int task = tl + BLK_HALF * m;
        if (task < n_task) {
            const int te = task >> 1, tn = tblock * Bs + te;
            double2 c = make_double2(0.0, 0.0), s = make_double2(0.0, 0.0);
            if (tn < N) {
                c = *reinterpret_cast<const double2*>((side ? R.x0b : R.x0k) + (size_t)tn * NQ + 2 * tp);
                s = *reinterpret_cast<const double2*>((side ? R.sum_b : R.sum_k) + (size_t)tn * NQ + 2 * tp);
            }
            *reinterpret_cast<double2*>(xsd + te * NQ + 2 * tp) = c;
            __stcg(reinterpret_cast<double2*>(my_sum + (size_t)te * NQ + 2 * tp), s);
            __stcg(reinterpret_cast<double2*>(my_prv + (size_t)te * NQ + 2 * tp), make_double2(0.0, 0.0));
        }
    }
    if (tid == 0) sctrl = *R.ctrl;

    // ---- product mappings
    // ket half: a thread owns rows o and o + Bh of the block; NGk thread groups split the columns of every chunk
    const int Bh = (Bs + 1) >> 1, Bo = (Bh + 31) & ~31;
    const int o = tl % Bo, gk = tl / Bo;
    const bool worker_k = (side == 0) && (gk < NGk) && (o < Bh);
    const bool validB = (o + Bh < Bs);
    const int oB = validB ? o + Bh : o;
    // bra half: a thread owns column cB of the chunk; NGb thread groups split the rows of the block
    const int cB = tl % Cc, gb = tl / Cc;
    const bool worker_b = (side == 1) && (gb < NGb);
    const int r0b = (Bs * gb) / NGb, r1b = (Bs * (gb + 1)) / NGb;

    // ---- chunk loader: a warp per column, lanes along the rows; chunk q of the series is chunk q % n_chunks of the block
    const int total_chunks = R.n_steps * n_chunks;                // <= 32 terms x (512 / 8) chunks
    auto issue_chunk = [&](int q) {
        const int kq = q % n_chunks, st = q % BLK_STAGES;
        const int c0 = kq * Cc, ncols = min(Cc, Bs - c0);
        double* dst0 = bsm + (size_t)st * L.stage_len;
        for (int cl = warp; cl < ncols; cl += BLK_THREADS / 32) {
            const int c = bj * Bs + c0 + cl;
            const double* src = R.H + (size_t)min(c, N - 1) * R.ld + (size_t)bi * Bs;
            for (int rl = lane; rl < Bs; rl += 32) {
                const bool ok = (c < N) && (bi * Bs + rl < N);
                cp_async_8(dst0 + cl * ldS + rl, ok ? src + rl : R.H, ok);
            }
        }
    };
    for (int q = 0; q < BLK_STAGES - 1; ++q) { if (q < total_chunks) issue_chunk(q); cp_async_commit(); }

    unsigned long long bar_target = 0;
    bool decided_all = false;
    double pass_word = 0.0;
    if (tid < sizeof(PassParams) / 8 && R.n_steps > 0) pass_word = reinterpret_cast<const double*>(R.passes)[tid];
    __syncthreads();

    int q = 0;
    int t = 0;
    for (; t < R.n_steps; ++t) {
        if (sctrl.part[0].latched && sctrl.part[1].latched) { decided_all = true; break; }
        if (tid < sizeof(PassParams) / 8) {
            reinterpret_cast<double*>(&spass[t & 1])[tid] = pass_word;
            if (t + 1 < R.n_steps) pass_word = reinterpret_cast<const double*>(R.passes + t + 1)[tid];
        }

        // ---------------------------------------------------------------- 1. stream the block, both products
        double aA[NQ] = {0.0, 0.0, 0.0, 0.0}, aB[NQ] = {0.0, 0.0, 0.0, 0.0};
        double* const pb_dst = R.pb + ((((size_t)(t & 1) * Gd + bi) * Gd + bj) * Bs) * NQ;     // bra partial of block column bj from block row bi
        for (int k = 0; k < n_chunks; ++k, ++q) {
            cp_async_wait<BLK_STAGES - 2>();
            __syncthreads();                                     // chunk q landed for everybody; everybody left chunk q-1
            if (q + BLK_STAGES - 1 < total_chunks) issue_chunk(q + BLK_STAGES - 1);
            cp_async_commit();
            const double* Hs = bsm + (size_t)(q % BLK_STAGES) * L.stage_len;
            const int c0 = k * Cc, ncols = min(Cc, Bs - c0);
            if (side == 0) {
                if (worker_k) {
                    const int j0 = (ncols * gk) / NGk, j1 = (ncols * (gk + 1)) / NGk;
#pragma unroll 2
                    for (int jc = j0; jc < j1; ++jc) {
                        const double h0 = Hs[jc * ldS + o], h1 = Hs[jc * ldS + oB];
                        const double2 x0 = *reinterpret_cast<const double2*>(xsd + (c0 + jc) * NQ), x1 = *reinterpret_cast<const double2*>(xsd + (c0 + jc) * NQ + 2);
                        aA[0] = fma(h0, x0.x, aA[0]); aA[1] = fma(h0, x0.y, aA[1]); aA[2] = fma(h0, x1.x, aA[2]); aA[3] = fma(h0, x1.y, aA[3]);
                        aB[0] = fma(h1, x0.x, aB[0]); aB[1] = fma(h1, x0.y, aB[1]); aB[2] = fma(h1, x1.x, aB[2]); aB[3] = fma(h1, x1.y, aB[3]);
                    }
                }
            } else {
                if (worker_b) {
                    double a0[NQ] = {0.0, 0.0, 0.0, 0.0}, a1[NQ] = {0.0, 0.0, 0.0, 0.0};
                    if (cB < ncols) {
                        const double* hc = Hs + cB * ldS;
                        int r = r0b;
                        for (; r + 1 < r1b; r += 2) {
                            const double h0 = hc[r], h1 = hc[r + 1];
                            const double2 x0 = *reinterpret_cast<const double2*>(xsd + r * NQ), x1 = *reinterpret_cast<const double2*>(xsd + r * NQ + 2);
                            const double2 y0 = *reinterpret_cast<const double2*>(xsd + (r + 1) * NQ), y1 = *reinterpret_cast<const double2*>(xsd + (r + 1) * NQ + 2);
                            a0[0] = fma(h0, x0.x, a0[0]); a0[1] = fma(h0, x0.y, a0[1]); a0[2] = fma(h0, x1.x, a0[2]); a0[3] = fma(h0, x1.y, a0[3]);
                            a1[0] = fma(h1, y0.x, a1[0]); a1[1] = fma(h1, y0.y, a1[1]); a1[2] = fma(h1, y1.x, a1[2]); a1[3] = fma(h1, y1.y, a1[3]);
                        }
                        if (r < r1b) {
                            const double h0 = hc[r];
                            const double2 x0 = *reinterpret_cast<const double2*>(xsd + r * NQ), x1 = *reinterpret_cast<const double2*>(xsd + r * NQ + 2);
                            a0[0] = fma(h0, x0.x, a0[0]); a0[1] = fma(h0, x0.y, a0[1]); a0[2] = fma(h0, x1.x, a0[2]); a0[3] = fma(h0, x1.y, a0[3]);
                        }
                    }
                    double2* pd = reinterpret_cast<double2*>(partb + (size_t)(gb * Cc + cB) * NQ);
                    pd[0] = make_double2(a0[0] + a1[0], a0[1] + a1[1]); pd[1] = make_double2(a0[2] + a1[2], a0[3] + a1[3]);
                }
                bar_sync_bra_half();                             // the bra half only: partb complete
                if (tl < 2 * ncols) {                            // (column, particle): fixed order over the row groups
                    const int c = tl >> 1;
                    double2 v = make_double2(0.0, 0.0);
                    for (int g = 0; g < NGb; ++g) {
                        const double2 p0 = *reinterpret_cast<const double2*>(partb + (size_t)(g * Cc + c) * NQ + 2 * tp);
                        v.x += p0.x; v.y += p0.y;
                    }
                    __stcg(reinterpret_cast<double2*>(pb_dst + (size_t)(c0 + c) * NQ + 2 * tp), v);
                }
                // partb is rewritten only after the next chunk's __syncthreads, which every reader above reaches first
            }
        }
        // ket partial of block row bi from block column bj -> pk[t&1][bj][bi][.]
        {
            double* const pk_dst = R.pk + ((((size_t)(t & 1) * Gd + bj) * Gd + bi) * Bs) * NQ;
            if (NGk > 1) {
                if (worker_k) {
                    double2* pA = reinterpret_cast<double2*>(partk + (size_t)(gk * Bs + o) * NQ);
                    pA[0] = make_double2(aA[0], aA[1]); pA[1] = make_double2(aA[2], aA[3]);
                    if (validB) {
                        double2* pB = reinterpret_cast<double2*>(partk + (size_t)(gk * Bs + oB) * NQ);
                        pB[0] = make_double2(aB[0], aB[1]); pB[1] = make_double2(aB[2], aB[3]);
                    }
                }
                __syncthreads();
                if (side == 0) {
                    for (int task = tl; task < n_task; task += BLK_HALF) {
                        const int e = task >> 1;
                        double2 v = make_double2(0.0, 0.0);
                        for (int g = 0; g < NGk; ++g) {
                            const double2 p0 = *reinterpret_cast<const double2*>(partk + (size_t)(g * Bs + e) * NQ + 2 * tp);
                            v.x += p0.x; v.y += p0.y;
                        }
                        __stcg(reinterpret_cast<double2*>(pk_dst + (size_t)e * NQ + 2 * tp), v);
                    }
                }
            } else if (worker_k) {
                __stcg(reinterpret_cast<double2*>(pk_dst + (size_t)o * NQ), make_double2(aA[0], aA[1]));
                __stcg(reinterpret_cast<double2*>(pk_dst + (size_t)o * NQ) + 1, make_double2(aA[2], aA[3]));
                if (validB) {
                    __stcg(reinterpret_cast<double2*>(pk_dst + (size_t)oB * NQ), make_double2(aB[0], aB[1]));
                    __stcg(reinterpret_cast<double2*>(pk_dst + (size_t)oB * NQ) + 1, make_double2(aB[2], aB[3]));
                }
            }
        }

        // ---------------------------------------------------------------- 2. the one grid barrier of the term
        bar_target += G;
        res_grid_barrier(R.gbar, bar_target);

        // ---------------------------------------------------------------- 3+4. per task: gather the 12 partials, then recurrence
        //                                  + series sum; the decision on term t-1 is taken while the first gather is in flight
        {
            const double* gsrc = (side ? R.pb : R.pk) + ((((size_t)(t & 1) * Gd) * Gd + tblock) * Bs) * NQ + 2 * tp;
            const size_t stride = (size_t)Gd * Bs * NQ;
            const PartPass& pa = spass[t & 1].part[tp];
            bool stop = false;
#pragma unroll 1
            for (int m = 0; m < BLK_MAX_TASK; ++m) {
                const int task = tl + BLK_HALF * m;
                const int te = task >> 1;
                const bool real = (task < n_task) && (tblock * Bs + te < N);
                double2 hx = make_double2(0.0, 0.0);
                if (real) {
                    const double* src = gsrc + (size_t)te * NQ;
                    double2 v[RES_MAX_GD];
#pragma unroll
                    for (int u = 0; u < RES_MAX_GD; ++u) if (u < Gd) v[u] = __ldcg(reinterpret_cast<const double2*>(src + u * stride));
#pragma unroll
                    for (int u = 0; u < RES_MAX_GD; ++u) if (u < Gd) { hx.x += v[u].x; hx.y += v[u].y; }
                }
                if (m == 0) {                                    // uniform: every thread runs BLK_MAX_TASK iterations
                    if (t > 0 && tid < 8) {
                        const double* ds = R.dscal + (size_t)((t + 1) & 1) * Gd * 8 + tid;
                        const bool is_max = (tid & 3) < 2;
                        double x[RES_MAX_GD];
#pragma unroll
                        for (int u = 0; u < RES_MAX_GD; ++u) x[u] = (u < Gd) ? __ldcg(ds + u * 8) : 0.0;
                        double f = 0.0;
#pragma unroll
                        for (int u = 0; u < RES_MAX_GD; ++u) f = is_max ? fmax(f, x[u]) : f + x[u];
                        fin[tid] = f;
                    }
                    __syncthreads();
                    if (t > 0) {
                        if (tid == 0 || tid == 32) { const int p = tid >> 5; decide_particle(sctrl.part[p], spass[(t - 1) & 1].part[p], fin + 4 * p); }
                        __syncthreads();
                        if (sctrl.part[0].latched && sctrl.part[1].latched) { stop = true; break; }
                    }
                }
                if (task >= n_task) continue;
                double mag = 0.0;
                if (real && pa.active && !sctrl.part[tp].latched) {
                    const double2 cur = *reinterpret_cast<const double2*>(xsd + te * NQ + 2 * tp);
                    const double2 sum = __ldcg(reinterpret_cast<const double2*>(my_sum + (size_t)te * NQ + 2 * tp));
                    Cx y = cmul({pa.alpha_re, pa.alpha_im}, {hx.x, hx.y});
                    if (pa.three_term) {
                        const Cx bc = cmul({pa.beta_re, pa.beta_im}, {cur.x, cur.y});
                        y.re += bc.re; y.im += bc.im;
                        if (pa.gamma != 0.0) {
                            const double2 prv = __ldcg(reinterpret_cast<const double2*>(my_prv + (size_t)te * NQ + 2 * tp));
                            y.re += pa.gamma * prv.x; y.im += pa.gamma * prv.y;
                        }
                        __stcg(reinterpret_cast<double2*>(my_prv + (size_t)te * NQ + 2 * tp), cur);
                    }
                    Cx tt = y;
                    if (pa.scale_term) tt = cmul({pa.c_re, pa.c_im}, y);
                    const double nw_re = sum.x + tt.re, nw_im = sum.y + tt.im;
                    const double dx = nw_re - sum.x, dy = nw_im - sum.y;
                    mag = dx * dx + dy * dy;                     // |new - old|^2 (isConverged, Taylor.f:290-303); root after the max
                    *reinterpret_cast<double2*>(xsd + te * NQ + 2 * tp) = make_double2(y.re, y.im);
                    __stcg(reinterpret_cast<double2*>(my_sum + (size_t)te * NQ + 2 * tp), make_double2(nw_re, nw_im));
                }
                if (diag) __stcg(my_mag + (size_t)te * 2 + tp, mag);
            }
            if (stop) { decided_all = true; break; }
        }
        __syncthreads();

        // ---------------------------------------------------------------- 5. diagonal CTAs: scalars of their block
        if (diag) {
            if (side == 0) {
                const double* sum_k_ = R.st_sum + ((size_t)(blockIdx.x * 2 + 0) * Bs) * NQ;
                const double* sum_b_ = R.st_sum + ((size_t)(blockIdx.x * 2 + 1) * Bs) * NQ;
                const double* mag_k_ = R.st_mag + ((size_t)(blockIdx.x * 2 + 0) * Bs) * 2;
                const double* mag_b_ = R.st_mag + ((size_t)(blockIdx.x * 2 + 1) * Bs) * 2;
                double v[4] = {0.0, 0.0, 0.0, 0.0};              // max_b, max_k, dot_re, dot_im of particle tp
                for (int task = tl; task < n_task; task += BLK_HALF) {
                    const int e = task >> 1;
                    const double2 k = __ldcg(reinterpret_cast<const double2*>(sum_k_ + (size_t)e * NQ + 2 * tp));
                    const double2 b = __ldcg(reinterpret_cast<const double2*>(sum_b_ + (size_t)e * NQ + 2 * tp));
                    v[0] = fmax(v[0], __ldcg(mag_b_ + (size_t)e * 2 + tp)); v[1] = fmax(v[1], __ldcg(mag_k_ + (size_t)e * 2 + tp));
                    v[2] += b.x * k.x + b.y * k.y;               // conj(bra) * ket
                    v[3] += b.x * k.y - b.y * k.x;
                }
#pragma unroll
                for (int off = 2; off < 32; off <<= 1) {         // lanes of equal particle (lane bit 0)
                    v[0] = fmax(v[0], __shfl_xor_sync(0xffffffffu, v[0], off)); v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], off));
                    v[2] += __shfl_xor_sync(0xffffffffu, v[2], off);            v[3] += __shfl_xor_sync(0xffffffffu, v[3], off);
                }
                if (lane < 2)
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) wred[warp][lane * 4 + qq] = v[qq];
            }
            __syncthreads();
            if (tid < 8) {
                double f = wred[0][tid];
                for (int w2 = 1; w2 < BLK_HALF / 32; ++w2) f = ((tid & 3) < 2) ? fmax(f, wred[w2][tid]) : f + wred[w2][tid];
                if ((tid & 3) < 2) f = sqrt(f);
                __stcg(R.dscal + ((size_t)(t & 1) * Gd + bi) * 8 + tid, f);
            }
        }
    }

    // ---- decision on the last term (one more barrier), unless the series was decided on the way
    if (!decided_all && t > 0) {
        bar_target += G;
        res_grid_barrier(R.gbar, bar_target);
        if (tid < 8) {
            const double* ds = R.dscal + (size_t)((t - 1) & 1) * Gd * 8 + tid;
            const bool is_max = (tid & 3) < 2;
            double v = 0.0;
            for (int u = 0; u < Gd; ++u) { const double x = __ldcg(ds + u * 8); v = is_max ? fmax(v, x) : v + x; }
            fin[tid] = v;
        }
        __syncthreads();
        if (tid == 0 || tid == 32) { const int p = tid >> 5; decide_particle(sctrl.part[p], spass[(t - 1) & 1].part[p], fin + 4 * p); }
        __syncthreads();
    }
    cp_async_wait<0>();                                          // no copy may land in a CTA that has left

    // ---- results: the diagonal CTAs hold the bra and ket sums of their block
    if (diag) {
        for (int task = tl; task < n_task; task += BLK_HALF) {
            const int te = task >> 1, tn = tblock * Bs + te;
            if (tn < N)
                *reinterpret_cast<double2*>((side ? R.sum_b : R.sum_k) + (size_t)tn * NQ + 2 * tp) =
                    __ldcg(reinterpret_cast<const double2*>(my_sum + (size_t)te * NQ + 2 * tp));
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        sctrl.all_latched = (sctrl.part[0].latched && sctrl.part[1].latched) ? 1 : 0;
        sctrl.block_counter = 0u;
        *R.ctrl = sctrl;
    }
}

}  // namespace dyb
