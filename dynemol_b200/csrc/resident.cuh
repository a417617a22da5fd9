// resident.cuh -- small operators (N <= ~1800): H' lives in SHARED MEMORY for a whole series.
//
// For the QM regions of the reference's Ehrenfest / CSDM examples (N of a few hundred to ~2000) a term of the series
// is not bandwidth work at all: H' (N^2 * 8 B <= 27 MB) fits in the shared memories of the 148 SMs taken together.
// One cooperative launch per series; a Gd x Gd grid of CTAs (12 x 12 on a B200), CTA (i, j) keeps block (i, j) of H'
// (Bs x Bs, Bs = ceil(N / Gd)) resident and, per term,
//
//   1. multiplies it from both sides:  ket partial  pk = H'(i,j) * x_ket[block j]   (Bs quads, owner = row)
//                                      bra partial  pb = H'(i,j)^T * x_bra[block i] (Bs quads, owner = column)
//      and stores both partials in L2 (double-buffered by the parity of the term),
//   2. ONE grid barrier,
//   3. rebuilds, redundantly, exactly the vector entries IT needs for the next term: x_ket on block j (sum of the Gd
//      ket partials of block row j) and x_bra on block i (sum of the Gd bra partials of block column i), applies the
//      recurrence and the series sum in registers (one thread per row and side, both particles).  Every CTA that
//      needs an entry computes it from the same partials in the same order, so all copies agree bit for bit.
//
// The convergence scalars come from the Gd diagonal CTAs (they hold bra and ket entries of the same block) and are
// consumed ONE TERM LATE: the decision on term t is taken by every CTA, identically, right after the barrier of
// term t+1 and before that term is added to the sums; a latched particle simply skips the update.  A series of L
// terms therefore costs L+1 grid barriers and no kernel boundary, and H' is read from HBM/L2 once per series.
//
// Same PassParams / Ctrl / apply_decision as the streaming path (epilogue.cuh): the host-side series logic
// (propagator.cu: propagate_series) does not know which kernel ran.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace dyb {

constexpr int RES_THREADS = 512;          // threads 0..255: ket side, 256..511: bra side
constexpr int RES_HALF    = 256;
constexpr int RES_MAX_BS  = 152;          // 152*153*8 B = 186 KB of H' per CTA
constexpr int RES_MAX_GD  = 18;          // multiple of the gather batch (6)
constexpr int RES_SMEM_MAX = 227 * 1024 - 2048;   // dynamic shared memory opt-in (the kernel's static part stays below 2 KB)

struct ResidentParams {
    const double* H; long long ld;        // column-major H' (N x N)
    int N, Gd, Bs, ldS;                   // grid side, block size, column stride of the block in shared memory (odd)
    const double* x0k; const double* x0b; // starting vectors (quads), written by series_init_kernel
    double* sum_b; double* sum_k;         // in: series sums at the start; out: at the latch / end of the series
    double* pk; double* pb;               // [2][Gd][Gd][Bs][NQ] partial products (parity of the term first)
    double* dscal;                        // [2][Gd][8] scalars of the diagonal CTAs
    Ctrl* ctrl;
    const PassParams* passes; int n_steps;
    unsigned long long* gbar;             // grid barrier counter (zeroed before the launch)
};

struct ResidentSmem {                     // dynamic shared memory carve-up (offsets in doubles)
    int hs, xk, xb, part, sq, total;
    __host__ __device__ ResidentSmem(int Bs, int ldS) {
        hs = 0;
        xk = (Bs * ldS + 1) & ~1;         // 16 B alignment for the quads
        xb = xk + Bs * NQ;
        part = xb + Bs * NQ;              // [2][RES_HALF][NQ] partials of the split inner range
        sq = part + 2 * RES_HALF * NQ;    // diagonal CTAs: updated sums [2][Bs][NQ] + term magnitudes [2][Bs][2]
        total = sq + 2 * Bs * NQ + 2 * Bs * 2;
    }
    __host__ __device__ size_t bytes() const { return (size_t)total * 8; }
};

__device__ __forceinline__ unsigned long long res_ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of the cooperative grid; bounded spin: a CTA that never arrives must end in a trap, not in a hung GPU
__device__ __forceinline__ void res_grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1ULL);
        for (long long it = 0; res_ld_acquire_gpu(ctr) < target; ++it)
            if (it > (1ll << 26)) __trap();
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RES_THREADS, 1)
resident_series_kernel(const ResidentParams R)
{
    extern __shared__ __align__(16) double rsm[];
    const ResidentSmem L(R.Bs, R.ldS);
    double* Hs = rsm + L.hs;
    double* xs[2] = {rsm + L.xk, rsm + L.xb};
    double* part = rsm + L.part;
    double* sq = rsm + L.sq;
    double* mq = sq + 2 * R.Bs * NQ;
    __shared__ Ctrl       sctrl;
    __shared__ PassParams spass[2];
    __shared__ double     fin[8];
    __shared__ double     wred[RES_HALF / 32][8];

    const int tid = threadIdx.x, lane = tid & 31;
    const int Gd = R.Gd, Bs = R.Bs, ldS = R.ldS, N = R.N;
    const int bi = blockIdx.x / Gd, bj = blockIdx.x % Gd;       // block row, block column
    const int G = Gd * Gd;
    const bool diag = (bi == bj);
    const int side = tid >> 8, tl = tid & (RES_HALF - 1);       // 0: ket (rows of block bi / entries of block bj), 1: bra

    // ---- H'(bi, bj) -> shared memory, column-major with an odd column stride (conflict-free from both sides)
    {
        const int r_base = bi * Bs, c_base = bj * Bs;
        for (int idx = tid; idx < Bs * Bs; idx += RES_THREADS) {
            const int cl = idx / Bs, rl = idx - cl * Bs;
            const int r = r_base + rl, c = c_base + cl;
            Hs[cl * ldS + rl] = (r < N && c < N) ? R.H[(size_t)c * R.ld + r] : 0.0;
        }
        if (tid == 0) sctrl = *R.ctrl;
    }

    // ---- epilogue task of this thread: entry n of block (side == 0 ? bj : bi), both particles, state in registers
    const int  tblock = side ? bi : bj;
    const int  tn = tblock * Bs + tl;                            // global index
    const bool task = (tl < Bs) && (tn < N);
    double cur[NQ] = {0.0, 0.0, 0.0, 0.0}, prv[NQ] = {0.0, 0.0, 0.0, 0.0}, sum[NQ] = {0.0, 0.0, 0.0, 0.0};
    if (task) {
        const double2* x0 = reinterpret_cast<const double2*>((side ? R.x0b : R.x0k) + (size_t)tn * NQ);
        const double2* s0 = reinterpret_cast<const double2*>((side ? R.sum_b : R.sum_k) + (size_t)tn * NQ);
        const double2 a = x0[0], b = x0[1], c = s0[0], d = s0[1];
        cur[0] = a.x; cur[1] = a.y; cur[2] = b.x; cur[3] = b.y;
        sum[0] = c.x; sum[1] = c.y; sum[2] = d.x; sum[3] = d.y;
    }
    if (tl < Bs) {
        double2* xd = reinterpret_cast<double2*>(xs[side] + tl * NQ);
        xd[0] = make_double2(cur[0], cur[1]); xd[1] = make_double2(cur[2], cur[3]);
    }

    // ---- product mapping: owner o (row for the ket side, column for the bra side), inner range split in NG groups
    const int Bo = (Bs + 31) & ~31;
    const int NG = RES_HALF / Bo;                               // >= 1 because Bs <= 152
    const int o = tl % Bo, g = tl / Bo;
    const bool worker = (g < NG) && (o < Bs);
    const int i0 = (Bs * g) / NG, i1 = (Bs * (g + 1)) / NG;
    const int so = side ? ldS : 1, si = side ? 1 : ldS;

    unsigned long long bar_target = 0;
    bool decided_all = false;
    __syncthreads();

    int t = 0;
    for (; t < R.n_steps; ++t) {
        if (sctrl.all_latched) { decided_all = true; break; }
        if (tid < sizeof(PassParams) / 8)
            reinterpret_cast<double*>(&spass[t & 1])[tid] = reinterpret_cast<const double*>(R.passes + t)[tid];

        // ---------------------------------------------------------------- 1. both products of the resident block
        {
            double a0[NQ] = {0.0, 0.0, 0.0, 0.0}, a1[NQ] = {0.0, 0.0, 0.0, 0.0};
            if (worker) {
                const double* hp = Hs + o * so;
                const double* xin = xs[side];
                int i = i0;
                for (; i + 1 < i1; i += 2) {                     // two independent chains per component
                    const double h0 = hp[i * si], h1 = hp[(i + 1) * si];
                    const double2 x0 = *reinterpret_cast<const double2*>(xin + i * NQ), x1 = *reinterpret_cast<const double2*>(xin + i * NQ + 2);
                    const double2 y0 = *reinterpret_cast<const double2*>(xin + (i + 1) * NQ), y1 = *reinterpret_cast<const double2*>(xin + (i + 1) * NQ + 2);
                    a0[0] = fma(h0, x0.x, a0[0]); a0[1] = fma(h0, x0.y, a0[1]); a0[2] = fma(h0, x1.x, a0[2]); a0[3] = fma(h0, x1.y, a0[3]);
                    a1[0] = fma(h1, y0.x, a1[0]); a1[1] = fma(h1, y0.y, a1[1]); a1[2] = fma(h1, y1.x, a1[2]); a1[3] = fma(h1, y1.y, a1[3]);
                }
                if (i < i1) {
                    const double h0 = hp[i * si];
                    const double2 x0 = *reinterpret_cast<const double2*>(xin + i * NQ), x1 = *reinterpret_cast<const double2*>(xin + i * NQ + 2);
                    a0[0] = fma(h0, x0.x, a0[0]); a0[1] = fma(h0, x0.y, a0[1]); a0[2] = fma(h0, x1.x, a0[2]); a0[3] = fma(h0, x1.y, a0[3]);
                }
            }
            double2* pd = reinterpret_cast<double2*>(part + (size_t)(side * RES_HALF + tl) * NQ);
            pd[0] = make_double2(a0[0] + a1[0], a0[1] + a1[1]); pd[1] = make_double2(a0[2] + a1[2], a0[3] + a1[3]);
        }
        __syncthreads();
        if (tl < Bs) {                                           // fixed order over the NG groups
            double v[NQ] = {0.0, 0.0, 0.0, 0.0};
            for (int gg = 0; gg < NG; ++gg) {
                const double2* ps = reinterpret_cast<const double2*>(part + (size_t)(side * RES_HALF + gg * Bo + tl) * NQ);
                const double2 p0 = ps[0], p1 = ps[1];
                v[0] += p0.x; v[1] += p0.y; v[2] += p1.x; v[3] += p1.y;
            }
            // ket partial of block row bi from block column bj -> pk[t&1][bj][bi][tl];  bra partial of block column bj
            // from block row bi -> pb[t&1][bi][bj][tl]
            double* dst = side ? R.pb + ((((size_t)(t & 1) * Gd + bi) * Gd + bj) * Bs + tl) * NQ
                               : R.pk + ((((size_t)(t & 1) * Gd + bj) * Gd + bi) * Bs + tl) * NQ;
            __stcg(reinterpret_cast<double2*>(dst), make_double2(v[0], v[1]));
            __stcg(reinterpret_cast<double2*>(dst) + 1, make_double2(v[2], v[3]));
        }

        // ---------------------------------------------------------------- 2. the one grid barrier of the term
        bar_target += G;
        res_grid_barrier(R.gbar, bar_target);

        // ---------------------------------------------------------------- 3. gather: partial sums for this thread's entry,
        //                                                                     and (one warp) the scalars of term t-1
        double hx[NQ] = {0.0, 0.0, 0.0, 0.0};
        if (task) {
            // ket entry n of block bj: sum over jj of pk[.][jj][bj][tl];  bra entry n of block bi: sum over ii of pb[.][ii][bi][tl]
            const double* src = (side ? R.pb : R.pk) + ((((size_t)(t & 1) * Gd) * Gd + tblock) * Bs + tl) * NQ;
            const size_t stride = (size_t)Gd * Bs * NQ;
#pragma unroll
            for (int u0 = 0; u0 < RES_MAX_GD; u0 += 6) {          // batches of 6 independent loads (register budget: 128)
                if (u0 >= Gd) break;
                double2 v0[6], v1[6];
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    if (u0 + u < Gd) { v0[u] = __ldcg(reinterpret_cast<const double2*>(src + (u0 + u) * stride)); v1[u] = __ldcg(reinterpret_cast<const double2*>(src + (u0 + u) * stride) + 1); }
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    if (u0 + u < Gd) { hx[0] += v0[u].x; hx[1] += v0[u].y; hx[2] += v1[u].x; hx[3] += v1[u].y; }
            }
        }
        if (t > 0 && tid >= RES_THREADS - 32 && lane < 8) {       // last warp never holds a task (Bs <= 152 < 224)
            const double* ds = R.dscal + (size_t)((t - 1) & 1) * Gd * 8 + lane;
            const bool is_max = (lane & 3) < 2;
            double x[RES_MAX_GD];
#pragma unroll
            for (int u = 0; u < RES_MAX_GD; ++u) x[u] = (u < Gd) ? __ldcg(ds + u * 8) : 0.0;
            double v = 0.0;
#pragma unroll
            for (int u = 0; u < RES_MAX_GD; ++u) v = is_max ? fmax(v, x[u]) : v + x[u];
            fin[lane] = v;
        }
        __syncthreads();
        if (t > 0) {
            if (tid == 0) apply_decision(&sctrl, spass[(t - 1) & 1], fin);
            __syncthreads();
            if (sctrl.all_latched) { decided_all = true; break; }
        }

        // ---------------------------------------------------------------- 4. recurrence + series sum (registers)
        double mag[2] = {0.0, 0.0};
        if (task) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const PartPass& pa = spass[t & 1].part[p];
                if (!pa.active || sctrl.part[p].latched) continue;
                Cx y = cmul({pa.alpha_re, pa.alpha_im}, {hx[2 * p], hx[2 * p + 1]});
                if (pa.three_term) {
                    const Cx bc = cmul({pa.beta_re, pa.beta_im}, {cur[2 * p], cur[2 * p + 1]});
                    y.re += bc.re; y.im += bc.im;
                    if (pa.gamma != 0.0) { y.re += pa.gamma * prv[2 * p]; y.im += pa.gamma * prv[2 * p + 1]; }
                }
                Cx tt = y;
                if (pa.scale_term) tt = cmul({pa.c_re, pa.c_im}, y);
                const double so_re = sum[2 * p], so_im = sum[2 * p + 1];
                const double nw_re = so_re + tt.re, nw_im = so_im + tt.im;
                mag[p] = hypot(nw_re - so_re, nw_im - so_im);             // |new - old| like isConverged (Taylor.f:290-303)
                prv[2 * p] = cur[2 * p]; prv[2 * p + 1] = cur[2 * p + 1];
                cur[2 * p] = y.re; cur[2 * p + 1] = y.im;
                sum[2 * p] = nw_re; sum[2 * p + 1] = nw_im;
            }
        }
        if (tl < Bs) {
            double2* xd = reinterpret_cast<double2*>(xs[side] + tl * NQ);
            xd[0] = make_double2(cur[0], cur[1]); xd[1] = make_double2(cur[2], cur[3]);
            if (diag) {
                double2* sd = reinterpret_cast<double2*>(sq + (size_t)(side * Bs + tl) * NQ);
                sd[0] = make_double2(sum[0], sum[1]); sd[1] = make_double2(sum[2], sum[3]);
                *reinterpret_cast<double2*>(mq + (size_t)(side * Bs + tl) * 2) = make_double2(mag[0], mag[1]);
            }
        }
        __syncthreads();

        // ---------------------------------------------------------------- 5. diagonal CTAs: scalars of their block
        if (diag) {
            if (side == 0) {                                     // 8 warps, entry tl of block bi == bj
                double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // {max_b, max_k, dot_re, dot_im} x particle
                if (tl < Bs) {
                    const double2* sk2 = reinterpret_cast<const double2*>(sq + (size_t)tl * NQ);
                    const double2* sb2 = reinterpret_cast<const double2*>(sq + (size_t)(Bs + tl) * NQ);
                    const double2 mk = *reinterpret_cast<const double2*>(mq + (size_t)tl * 2);
                    const double2 mb = *reinterpret_cast<const double2*>(mq + (size_t)(Bs + tl) * 2);
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const double2 k = sk2[p], b = sb2[p];
                        v[p * 4 + 0] = p ? mb.y : mb.x; v[p * 4 + 1] = p ? mk.y : mk.x;
                        v[p * 4 + 2] = b.x * k.x + b.y * k.y;    // conj(bra) * ket
                        v[p * 4 + 3] = b.x * k.y - b.y * k.x;
                    }
                }
#pragma unroll
                for (int off = 1; off < 32; off <<= 1)
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double ov = __shfl_xor_sync(0xffffffffu, v[q], off);
                        v[q] = ((q & 3) < 2) ? fmax(v[q], ov) : v[q] + ov;
                    }
                if (lane == 0)
#pragma unroll
                    for (int q = 0; q < 8; ++q) wred[tl >> 5][q] = v[q];
            }
            __syncthreads();
            if (tid < 8) {
                double f = wred[0][tid];
                for (int w2 = 1; w2 < RES_HALF / 32; ++w2) f = ((tid & 3) < 2) ? fmax(f, wred[w2][tid]) : f + wred[w2][tid];
                __stcg(R.dscal + ((size_t)(t & 1) * Gd + bi) * 8 + tid, f);
            }
        }
    }

    // ---- decision on the last term (one more barrier), unless the series was decided on the way
    if (!decided_all && t > 0) {
        bar_target += G;
        res_grid_barrier(R.gbar, bar_target);
        if (tid >= RES_THREADS - 32 && lane < 8) {
            const double* ds = R.dscal + (size_t)((t - 1) & 1) * Gd * 8 + lane;
            const bool is_max = (lane & 3) < 2;
            double v = 0.0;
            for (int u = 0; u < Gd; ++u) { const double x = __ldcg(ds + u * 8); v = is_max ? fmax(v, x) : v + x; }
            fin[lane] = v;
        }
        __syncthreads();
        if (tid == 0) apply_decision(&sctrl, spass[(t - 1) & 1], fin);
        __syncthreads();
    }

    // ---- results: the diagonal CTAs hold the bra and ket sums of their block
    if (diag && task) {
        double2* d = reinterpret_cast<double2*>((side ? R.sum_b : R.sum_k) + (size_t)tn * NQ);
        d[0] = make_double2(sum[0], sum[1]); d[1] = make_double2(sum[2], sum[3]);
    }
    if (blockIdx.x == 0 && tid == 0) { const unsigned keep = R.ctrl->block_counter; *R.ctrl = sctrl; R.ctrl->block_counter = keep; }
}

}  // namespace dyb
