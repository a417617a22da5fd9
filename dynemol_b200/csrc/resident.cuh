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

constexpr int RES_THREADS = 640;          // threads 0..319: ket side, 320..639: bra side
constexpr int RES_HALF    = 320;
constexpr int RES_MAX_BS  = 152;          // 152*153*8 B = 186 KB of H' per CTA
constexpr int RES_MAX_GD  = 12;          // grid side (12 x 12 = 144 CTAs on the 148 SMs of a B200)
constexpr int RES_NS      = EPI_NS_RG;   // scalar slots per diagonal CTA (8 + 4 arg-max keys, used by the parity modes only)
constexpr int RES_SMEM_MAX = 227 * 1024 - 2048;   // dynamic shared memory opt-in (the kernel's static part stays below 2 KB)

struct ResidentParams {
    const double* H; long long ld;        // column-major H' (N x N)
    int N, Gd, Bs, ldS;                   // grid side, block size, column stride of the block in shared memory (odd)
    const double* x0k; const double* x0b; // starting vectors (quads), written by series_init_kernel
    double* sum_b; double* sum_k;         // in: series sums at the start; out: at the latch / end of the series
    double* pk; double* pb;               // [2][Gd][Gd][Bs][NQ] partial products (parity of the term first)
    double* dscal;                        // [2][Gd][RES_NS] scalars of the diagonal CTAs (slots: epilogue.cuh EPI_NS_RG)
    double2* psi_store;                   // [grid][RES_THREADS] start vector of the sub-step in progress (needed only after a failure)
    Ctrl* ctrl;
    const PassParams* passes; int n_steps;
    unsigned long long* gbar;             // grid barrier counter (zeroed before the launch)
    long long* prof;                      // DYB_SERIES_PROF builds: [n_steps][grid][6] clock64 stamps (else null)
};

#ifdef DYB_SERIES_PROF
#define DYB_RSTAMP(i) do { if (threadIdx.x == 0 && t < 32) R.prof[((size_t)t * G + blockIdx.x) * 6 + (i)] = clock64(); } while (0)
#else
#define DYB_RSTAMP(i) do { } while (0)
#endif

struct ResidentSmem {                     // dynamic shared memory carve-up (offsets in doubles)
    int hs, xk, xb, part, sq, total;
    __host__ __device__ ResidentSmem(int Bs, int ldS) {
        hs = 0;
        xk = (Bs * ldS + 1) & ~1;         // 16 B alignment for the quads
        xb = xk + Bs * NQ;
        part = xb + Bs * NQ;              // [2][NG][Bs][NQ] partials of the NG inner-range groups
        sq = part;                        // diagonal CTAs: updated sums [2][Bs][NQ] + term magnitudes [2][Bs][2]; shares the
                                          // space of `part` (dead between the partial store and the next product)
        const int Bh = (Bs + 1) >> 1, NG = RES_HALF / ((Bh + 31) & ~31);
        total = part + 2 * NG * Bs * NQ;
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
        // release-add orders the CTA's writes (made visible to this thread by the __syncthreads above) before the
        // arrival; the acquire poll orders the other CTAs' writes before everything after the second __syncthreads
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
        for (long long it = 0; res_ld_acquire_gpu(ctr) < target; ++it)
            if (it > (1ll << 26)) __trap();
    }
    __syncthreads();
}

template <bool RG>        // RG: reference-GPU term test (parity modes); a separate instantiation keeps the default one lean
__global__ void __launch_bounds__(RES_THREADS, 1)
resident_series_kernel_t(const ResidentParams R)
{
    extern __shared__ __align__(16) double rsm[];
    const ResidentSmem L(R.Bs, R.ldS);
    double* Hs = rsm + L.hs;
    double* part = rsm + L.part;
    double* sq = rsm + L.sq;
    double* mq = sq + 2 * R.Bs * NQ;
    double* kq = mq + 2 * R.Bs * 2;                              // refgpu: max(re^2, im^2) of new - old (the Idamax key, squared)
    __shared__ Ctrl       sctrl;
    __shared__ PassParams spass[2];
    __shared__ double     fin[RES_NS];
    __shared__ double     wred[RES_HALF / 32][RES_NS];
    __shared__ int        stop_chain;                            // a particle failed a chained sub-step: the other one stops at
                                                                 // its next sub-step boundary so that both resume together

    const int tid = threadIdx.x, lane = tid & 31;
    const int Gd = R.Gd, Bs = R.Bs, ldS = R.ldS, N = R.N;
    const int bi = blockIdx.x / Gd, bj = blockIdx.x % Gd;       // block row, block column
    const int G = Gd * Gd;
    const bool diag = (bi == bj);
    const int side = tid / RES_HALF, tl = tid - side * RES_HALF; // 0: ket (rows of block bi / entries of block bj), 1: bra
    double* const xsd = rsm + (side ? L.xb : L.xk);             // this half's input vector block (quads) in shared memory

    // ---- H'(bi, bj) -> shared memory, column-major with an odd column stride (conflict-free from both sides)
    {
        const int r_base = bi * Bs, c_base = bj * Bs;
        // a warp per column, lanes along the rows (coalesced), up to 20 independent loads in flight per thread
#pragma unroll 4
        for (int cl = tid >> 5; cl < Bs; cl += RES_THREADS / 32) {
            const int c = c_base + cl;
            const double* src = R.H + (size_t)c * R.ld + r_base;
#pragma unroll
            for (int k = 0; k < (RES_MAX_BS + 31) / 32; ++k) {
                const int rl = lane + 32 * k;
                if (rl < Bs) Hs[cl * ldS + rl] = (r_base + rl < N && c < N) ? __ldg(src + rl) : 0.0;
            }
        }
        if (tid == 0) { sctrl = *R.ctrl; stop_chain = 0; }
    }

    // ---- epilogue task of this thread: particle tp of entry te of block (side == 0 ? bj : bi); state in registers
    const int  tblock = side ? bi : bj;
    const int  te = tl >> 1, tp = tl & 1;
    const int  tn = tblock * Bs + te;                            // global index
    const bool slot = (te < Bs);                                 // owns a (possibly padding) entry of the block
    const bool task = slot && (tn < N);
    double2 cur = make_double2(0.0, 0.0), prv = make_double2(0.0, 0.0), sum = make_double2(0.0, 0.0);
    // start vector of the sub-step in progress (chained sub-steps): only a failed norm test needs it back, so it is
    // parked in global memory (one fire-and-forget store per sub-step) instead of occupying registers
    double2* const psi_slot = R.psi_store + (size_t)blockIdx.x * RES_THREADS + tid;
    if (task) {
        cur = *reinterpret_cast<const double2*>((side ? R.x0b : R.x0k) + (size_t)tn * NQ + 2 * tp);
        sum = *reinterpret_cast<const double2*>((side ? R.sum_b : R.sum_k) + (size_t)tn * NQ + 2 * tp);
        __stcg(psi_slot, cur);
    }
    bool was_active = false;                                     // set below from the control block: only particles that take
                                                                 // part in this launch hand a vector back
    if (slot) *reinterpret_cast<double2*>(xsd + te * NQ + 2 * tp) = cur;

    // ---- product mapping: a thread owns TWO owners oA = o and oB = o + Bh (rows for the ket side, columns for the bra
    // side) so that every broadcast x quad read from shared memory feeds 8 FMAs; the inner range is split in NG groups
    const int Bh = (Bs + 1) >> 1;
    const int Bo = (Bh + 31) & ~31;
    const int NG = RES_HALF / Bo;                               // >= 3 because Bs <= 152
    const int o = tl % Bo, g = tl / Bo;
    const bool worker = (g < NG) && (o < Bh);
    const bool validB = (o + Bh < Bs);
    const int i0 = (Bs * g) / NG, i1 = (Bs * (g + 1)) / NG;
    const int so = side ? ldS : 1, si = side ? 1 : ldS;

    unsigned long long bar_target = 0;
    bool decided_all = false;
    double pass_word = 0.0;
    if (tid < sizeof(PassParams) / 8 && R.n_steps > 0) pass_word = reinterpret_cast<const double*>(R.passes)[tid];
    __syncthreads();
    was_active = !sctrl.part[tp].latched;                        // series_init_kernel latches the particles that sit out

    int t = 0;
    for (; t < R.n_steps; ++t) {
        if (sctrl.part[0].latched && sctrl.part[1].latched) { decided_all = true; break; }
        if (tid < sizeof(PassParams) / 8) {                      // this term's parameters were fetched one term ahead
            reinterpret_cast<double*>(&spass[t & 1])[tid] = pass_word;
            if (t + 1 < R.n_steps) pass_word = reinterpret_cast<const double*>(R.passes + t + 1)[tid];
        }

        DYB_RSTAMP(0);
        // ---------------------------------------------------------------- 1. both products of the resident block
        if (worker) {
            double aA[NQ] = {0.0, 0.0, 0.0, 0.0}, aB[NQ] = {0.0, 0.0, 0.0, 0.0};
            const double* hA = Hs + o * so;
            const double* hB = Hs + (validB ? o + Bh : o) * so;
#pragma unroll 2
            for (int i = i0; i < i1; ++i) {
                const double h0 = hA[i * si], h1 = hB[i * si];
                const double2 x0 = *reinterpret_cast<const double2*>(xsd + i * NQ), x1 = *reinterpret_cast<const double2*>(xsd + i * NQ + 2);
                aA[0] = fma(h0, x0.x, aA[0]); aA[1] = fma(h0, x0.y, aA[1]); aA[2] = fma(h0, x1.x, aA[2]); aA[3] = fma(h0, x1.y, aA[3]);
                aB[0] = fma(h1, x0.x, aB[0]); aB[1] = fma(h1, x0.y, aB[1]); aB[2] = fma(h1, x1.x, aB[2]); aB[3] = fma(h1, x1.y, aB[3]);
            }
            double2* pA = reinterpret_cast<double2*>(part + ((size_t)(side * NG + g) * Bs + o) * NQ);
            pA[0] = make_double2(aA[0], aA[1]); pA[1] = make_double2(aA[2], aA[3]);
            if (validB) {
                double2* pB = reinterpret_cast<double2*>(part + ((size_t)(side * NG + g) * Bs + o + Bh) * NQ);
                pB[0] = make_double2(aB[0], aB[1]); pB[1] = make_double2(aB[2], aB[3]);
            }
        }
        __syncthreads();
        DYB_RSTAMP(1);
        if (slot) {                                              // fixed order over the NG groups; this thread: one particle
            double2 v = make_double2(0.0, 0.0);
#pragma unroll 5
            for (int gg = 0; gg < NG; ++gg) {
                const double2 p0 = *reinterpret_cast<const double2*>(part + ((size_t)(side * NG + gg) * Bs + te) * NQ + 2 * tp);
                v.x += p0.x; v.y += p0.y;
            }
            // ket partial of block row bi from block column bj -> pk[t&1][bj][bi][te];  bra partial of block column bj
            // from block row bi -> pb[t&1][bi][bj][te]
            double* dst = side ? R.pb + ((((size_t)(t & 1) * Gd + bi) * Gd + bj) * Bs + te) * NQ
                               : R.pk + ((((size_t)(t & 1) * Gd + bj) * Gd + bi) * Bs + te) * NQ;
            __stcg(reinterpret_cast<double2*>(dst + 2 * tp), v);
        }

        DYB_RSTAMP(2);
        // ---------------------------------------------------------------- 2. the one grid barrier of the term
        bar_target += G;
        res_grid_barrier(R.gbar, bar_target);

        DYB_RSTAMP(3);
        // ---------------------------------------------------------------- 3. gather: partial sums for this thread's entry,
        //                                                                     and (eight threads) the scalars of term t-1
        double2 hx = make_double2(0.0, 0.0);
        {
            // Task threads: ket entry of block bj = sum over jj of pk[.][jj][bj][te]; bra entry of block bi = sum over ii
            // of pb[.][ii][bi][te].  The last four threads (never a task: Bs <= 152) fetch, the same way, the scalars the
            // diagonal CTAs left for term t-1: thread j takes components 2j, 2j+1 = a pair of maxima or a pair of sums.
            const bool scal = (t > 0) && (tid >= RES_THREADS - 4);
            const int  j2 = 2 * (tid - (RES_THREADS - 4));
            const double* src = scal ? R.dscal + (size_t)((t + 1) & 1) * Gd * RES_NS + j2
                                     : (side ? R.pb : R.pk) + ((((size_t)(t & 1) * Gd) * Gd + tblock) * Bs + te) * NQ + 2 * tp;
            const size_t stride = scal ? (size_t)RES_NS : (size_t)Gd * Bs * NQ;
            if (task || scal) {
                double2 v[RES_MAX_GD];                           // all Gd values in flight at once: one L2 round trip
#pragma unroll
                for (int u = 0; u < RES_MAX_GD; ++u) if (u < Gd) v[u] = __ldcg(reinterpret_cast<const double2*>(src + u * stride));
                if (RG && scal && (j2 & 2) == 0) {         // arg-max with payload: keys in slots 8 + j2/2, 9 + j2/2
                    const double* ksrc = R.dscal + (size_t)((t + 1) & 1) * Gd * RES_NS + 8 + (j2 >> 1);
                    double2 key = make_double2(0.0, 0.0);
                    for (int u = 0; u < Gd; ++u) {
                        const double2 ku = __ldcg(reinterpret_cast<const double2*>(ksrc + u * stride));
                        argmax_merge(key.x, hx.x, ku.x, v[u].x); argmax_merge(key.y, hx.y, ku.y, v[u].y);
                    }
                } else if (scal && (j2 & 2) == 0) {              // maxima (of non-negative numbers)
#pragma unroll
                    for (int u = 0; u < RES_MAX_GD; ++u) if (u < Gd) { hx.x = fmax(hx.x, v[u].x); hx.y = fmax(hx.y, v[u].y); }
                } else {
#pragma unroll
                    for (int u = 0; u < RES_MAX_GD; ++u) if (u < Gd) { hx.x += v[u].x; hx.y += v[u].y; }
                }
                if (scal) { fin[j2] = hx.x; fin[j2 + 1] = hx.y; }
            }
        }
        __syncthreads();
        if (t > 0) {                                             // one thread per particle, in different warps
            if (tid == 0 || tid == 32) {
                const int p = tid >> 5;                          // stop_chain as of the previous term: the same in every CTA
                decide_particle(sctrl.part[p], spass[(t - 1) & 1].part[p], fin + 4 * p, !stop_chain);
            }
            __syncthreads();
            if (sctrl.part[0].latched && sctrl.part[1].latched) { decided_all = true; break; }
            if (tid == 0 && ((sctrl.part[0].latched && !sctrl.part[0].ok) || (sctrl.part[1].latched && !sctrl.part[1].ok))) stop_chain = 1;
        }

        DYB_RSTAMP(4);
        // ---------------------------------------------------------------- 4. recurrence + series sum (registers)
        double mag = 0.0, key = 0.0;
        double2 xout = cur;                                      // what the next product multiplies
        if (task) {
            const PartPass& pa = spass[t & 1].part[tp];
            if (pa.active && !sctrl.part[tp].latched) {
                if (pa.begin) {                                  // next steady sub-step: adopt the previous sum (Taylor.f:105,
                    __stcg(psi_slot, sum); cur = sum;            // :83-86); hx was computed from it (xout of the last term)
                    const Cx s0 = cmul({pa.s_re, pa.s_im}, {sum.x, sum.y});
                    sum = make_double2(s0.re, s0.im);
                }
                Cx y = cmul({pa.alpha_re, pa.alpha_im}, {hx.x, hx.y});
                if (pa.three_term) {
                    const Cx bc = cmul({pa.beta_re, pa.beta_im}, {cur.x, cur.y});
                    y.re += bc.re; y.im += bc.im;
                    if (pa.gamma != 0.0) { y.re += pa.gamma * prv.x; y.im += pa.gamma * prv.y; }
                }
                Cx tt = y;
                if (pa.scale_term) tt = cmul({pa.c_re, pa.c_im}, y);
                const double nw_re = sum.x + tt.re, nw_im = sum.y + tt.im;
                const double dx = nw_re - sum.x, dy = nw_im - sum.y;
                mag = dx * dx + dy * dy;                         // |new - old|^2 (isConverged, Taylor.f:290-303); root after the max
                key = fmax(dx * dx, dy * dy);                    // what cublasIdamax ranks (Taylor_gpu.cpp:84-89), squared
                prv = cur;
                cur = make_double2(y.re, y.im);
                sum = make_double2(nw_re, nw_im);
                xout = pa.chain ? sum : cur;                     // speculative: the next sub-step starts from this sum
            }
        }
        if (slot) {
            *reinterpret_cast<double2*>(xsd + te * NQ + 2 * tp) = xout;
            if (diag) {
                *reinterpret_cast<double2*>(sq + (size_t)(side * Bs + te) * NQ + 2 * tp) = sum;
                mq[(size_t)(side * Bs + te) * 2 + tp] = mag;
                if constexpr (RG) kq[(size_t)(side * Bs + te) * 2 + tp] = key;
            }
        }
        __syncthreads();

        // ---------------------------------------------------------------- 5. diagonal CTAs: scalars of their block
        if (diag) {
            if (side == 0) {                                     // 10 warps; thread = (entry te, particle tp) of block bi == bj
                double v[4] = {0.0, 0.0, 0.0, 0.0};              // max_b, max_k, dot_re, dot_im of this particle
                double kv[2] = {0.0, 0.0};                       // refgpu: arg-max keys of bra, ket
                if (slot) {
                    const double2 k = *reinterpret_cast<const double2*>(sq + (size_t)te * NQ + 2 * tp);
                    const double2 b = *reinterpret_cast<const double2*>(sq + (size_t)(Bs + te) * NQ + 2 * tp);
                    v[0] = mq[(size_t)(Bs + te) * 2 + tp]; v[1] = mq[(size_t)te * 2 + tp];
                    if constexpr (RG) { kv[0] = kq[(size_t)(Bs + te) * 2 + tp]; kv[1] = kq[(size_t)te * 2 + tp]; }
                    v[2] = b.x * k.x + b.y * k.y;                // conj(bra) * ket
                    v[3] = b.x * k.y - b.y * k.x;
                }
#pragma unroll
                for (int off = 2; off < 32; off <<= 1) {         // lanes of equal particle (lane bit 0)
                    if constexpr (RG) {
                        const double o0 = __shfl_xor_sync(0xffffffffu, v[0], off), k0 = __shfl_xor_sync(0xffffffffu, kv[0], off);
                        const double o1 = __shfl_xor_sync(0xffffffffu, v[1], off), k1 = __shfl_xor_sync(0xffffffffu, kv[1], off);
                        argmax_merge(kv[0], v[0], k0, o0); argmax_merge(kv[1], v[1], k1, o1);
                    } else {
                        v[0] = fmax(v[0], __shfl_xor_sync(0xffffffffu, v[0], off)); v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], off));
                    }
                    v[2] += __shfl_xor_sync(0xffffffffu, v[2], off);            v[3] += __shfl_xor_sync(0xffffffffu, v[3], off);
                }
                if (lane < 2) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) wred[tl >> 5][lane * 4 + q] = v[q];
                    if constexpr (RG) { wred[tl >> 5][8 + 2 * lane] = kv[0]; wred[tl >> 5][9 + 2 * lane] = kv[1]; }
                }
            }
            __syncthreads();
            if constexpr (RG) {
                if (tid == 0) {                                  // parity mode: one thread merges the warps in order
                    double f[RES_NS];
#pragma unroll
                    for (int q = 0; q < RES_NS; ++q) f[q] = wred[0][q];
                    for (int w2 = 1; w2 < RES_HALF / 32; ++w2) scal_merge<RES_NS>(f, wred[w2]);
#pragma unroll
                    for (int q = 0; q < RES_NS; ++q)
                        __stcg(R.dscal + ((size_t)(t & 1) * Gd + bi) * RES_NS + q, (q < 8 && (q & 3) < 2) ? sqrt(f[q]) : f[q]);
                }
            } else if (tid < 8) {
                double wv[RES_HALF / 32];                        // all loads first: one shared-memory latency, not ten
#pragma unroll
                for (int w2 = 0; w2 < RES_HALF / 32; ++w2) wv[w2] = wred[w2][tid];
                double f = wv[0];
#pragma unroll
                for (int w2 = 1; w2 < RES_HALF / 32; ++w2) f = ((tid & 3) < 2) ? fmax(f, wv[w2]) : f + wv[w2];
                if ((tid & 3) < 2) f = sqrt(f);
                __stcg(R.dscal + ((size_t)(t & 1) * Gd + bi) * RES_NS + tid, f);
            }
        }
        DYB_RSTAMP(5);
    }

    // ---- decision on the last term (one more barrier), unless the series was decided on the way
    if (!decided_all && t > 0) {
        bar_target += G;
        res_grid_barrier(R.gbar, bar_target);
        if (tid >= RES_THREADS - 8) {
            const int q = tid & 7;
            const double* ds = R.dscal + (size_t)((t - 1) & 1) * Gd * RES_NS + q;
            const bool is_max = (q & 3) < 2;
            double v = 0.0, kbest = 0.0;
            for (int u = 0; u < Gd; ++u) {
                const double x = __ldcg(ds + u * RES_NS);
                if (RG && is_max) argmax_merge(kbest, v, __ldcg(R.dscal + ((size_t)((t - 1) & 1) * Gd + u) * RES_NS + 8 + 2 * (q >> 2) + (q & 1)), x);
                else v = is_max ? fmax(v, x) : v + x;
            }
            fin[q] = v;
        }
        __syncthreads();
        if (tid == 0 || tid == 32) { const int p = tid >> 5; decide_particle(sctrl.part[p], spass[(t - 1) & 1].part[p], fin + 4 * p); }
        __syncthreads();
    }

    // ---- results: the diagonal CTAs hold the bra and ket sums of their block
    // a particle that failed a steady sub-step hands back the start vector of that sub-step (= the last accepted sum)
    if (diag && task && was_active)
        *reinterpret_cast<double2*>((side ? R.sum_b : R.sum_k) + (size_t)tn * NQ + 2 * tp) =
            (sctrl.part[tp].latched && !sctrl.part[tp].ok) ? __ldcg(psi_slot) : sum;
    if (blockIdx.x == 0 && tid == 0) {
        sctrl.all_latched = (sctrl.part[0].latched && sctrl.part[1].latched) ? 1 : 0;
        sctrl.block_counter = 0u;
        *R.ctrl = sctrl;
    }
}

}  // namespace dyb
