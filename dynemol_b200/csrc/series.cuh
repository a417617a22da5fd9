// series.cuh -- ONE cooperative launch per series: every term's dual product AND its epilogue run inside the same
// persistent kernel (one CTA per SM), separated by grid barriers instead of kernel boundaries.
//
//   term t:  [tile loop over H' (TMA ring, as in matvec.cuh)] -> grid barrier A ->
//            [epilogue on this CTA's share of the rows: slab sums, recurrence, series sum, 8 scalars] -> grid barrier B ->
//            [every CTA combines the per-CTA scalars in CTA order and takes the same decision]
//
// The TMA ring never drains between terms: while the CTAs sit in the barriers and the epilogue, the first
// TMA_STAGES tiles of term t+1 are already in flight (their H' part does not depend on the epilogue; the x_ket part is
// issued right after barrier B).  Per term this replaces two kernel launches and the pipeline refill of the
// launch-per-term path; for small operators it removes the launch latency that dominates there.
//
// Memory visibility inside one launch: vectors and slabs written by other CTAs are read with ld.global.cg after a
// grid barrier (atomic arrive + acquire spin, __threadfence on both sides); the x_ket bulk copies (async proxy) are
// issued after a fence.proxy.async by the issuing thread.
#pragma once
#include "matvec.cuh"
#include "epilogue.cuh"

namespace dyb {

#ifdef DYB_SERIES_PROF
#define DYB_STAMP(i) do { if (threadIdx.x == 0) S.prof[((size_t)t * G + b) * 6 + (i)] = clock64(); } while (0)
#else
#define DYB_STAMP(i) do { } while (0)
#endif

struct SeriesParams {
    MatvecParams mv;                 // H', plan, slabs (Xk/Xb/ctrl members unused here)
    int row0, n_bra_slabs;           // epilogue: owned rows start at global index row0 (single GPU: 0)
    const int* pseg_start;
    double* vb[3]; double* vk[3];    // rotating vectors (prv, cur, nxt start as indices 2, 0, 1)
    double* sum_b; double* sum_k;
    double* blockpart;               // [grid][8]
    Ctrl*   ctrl;
    const PassParams* passes;        // [n_steps] per-term parameters (device memory)
    int n_steps;
    unsigned long long* gbar;        // grid barrier counter, zeroed by the host before the launch
    long long* prof;                 // DYB_SERIES_PROF builds: [n_steps][grid][6] clock64 stamps of the phases (else null)
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the (cooperative) grid; `target` = number of arrivals that completes this barrier
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1ULL);
        for (long long it = 0; ld_acquire_gpu(ctr) < target; ++it) {
            if (it > (1ll << 26)) __trap();           // a CTA that never arrives must not hang the GPU
        }
        __threadfence();
    }
    __syncthreads();
}

// bounded mbarrier wait: a protocol bug must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    for (long long it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > (1ll << 26)) __trap();
}

__device__ __forceinline__ void load_xb_coherent(ConsumerRegs& r, const double* Xb, long long panel, int w, int lane) {
    const double2* base = reinterpret_cast<const double2*>(Xb + ((panel * PANEL_ROWS + w * SUB_ROWS + 2 * lane) * NQ));
#pragma unroll
    for (int m = 0; m < MPT; ++m) {
        const double2* p = base + m * (64 * NQ / 2);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double2 v0 = __ldcg(p + e * 2), v1 = __ldcg(p + e * 2 + 1);     // written by other CTAs in this launch
            r.xb[m][e][0] = v0.x; r.xb[m][e][1] = v0.y; r.xb[m][e][2] = v1.x; r.xb[m][e][3] = v1.y;
        }
    }
}

// rotating vector `idx` (0..2) without dynamic indexing into the kernel parameters (that would force a local copy)
__device__ __forceinline__ double* pick3(double* const (&v)[3], int idx) { return idx == 0 ? v[0] : (idx == 1 ? v[1] : v[2]); }

// Epilogue of one term on rows r0..r1-1 (this CTA's share): slab sums, recurrence, series sum, and the CTA's 8 scalars
// into S.blockpart[blockIdx.x].  Kept out of line: its registers must not extend the live ranges of the tile loop.
__device__ __noinline__ void epilogue_phase(const SeriesParams& S, const PassParams& pass, const Ctrl& sctrl,
                                            int r0, int r1, int prv, int cur, int nxt, double (*wpart)[8])
{
    const MatvecParams& P = S.mv;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.x;
        double mb = 0.0, mk = 0.0, dr = 0.0, di = 0.0;      // per-thread partials, particle = (threadIdx.x >> 1) & 1
        for (int base = 0; base < (r1 - r0) * 4; base += TMA_THREADS) {
            const int task = base + threadIdx.x;
            const int i = r0 + (task >> 2), pp = (task >> 1) & 1, side = task & 1;
            const PartPass pa = pass.part[pp];
            const bool live = (i < r1) && pa.active && !sctrl.part[pp].latched;
            double mx = 0.0;
            Cx nw = {0.0, 0.0};
            if (live) {
                const size_t ob = (size_t)i * NQ + 2 * pp;
                const size_t og = ((size_t)S.row0 + i) * NQ + 2 * pp;
                Cx hx;
                if (side == 0) {
                    const int panel = i / PANEL_ROWS, il = i % PANEL_ROWS;
                    const int q0 = S.pseg_start[panel], q1 = S.pseg_start[panel + 1];
                    hx = slab_sum(P.ket_slab + (((size_t)q0 * PANEL_ROWS + il) * NQ + 2 * pp), (size_t)PANEL_ROWS * NQ, q1 - q0);
                } else {
                    hx = slab_sum(P.bra_slab + (((size_t)S.row0 + i) * NQ + 2 * pp), (size_t)P.Ncpad * NQ, S.n_bra_slabs);
                }
                Cx y = cmul({pa.alpha_re, pa.alpha_im}, hx);
                const size_t ov = side ? ob : og;
                const double* xc = side ? pick3(S.vb, cur) : pick3(S.vk, cur);
                const double* xp = side ? pick3(S.vb, prv) : pick3(S.vk, prv);
                double* xn = side ? pick3(S.vb, nxt) : pick3(S.vk, nxt);
                double* sum = side ? S.sum_b : S.sum_k;
                if (pa.three_term) {
                    const double2 c = __ldcg(reinterpret_cast<const double2*>(xc + ov));
                    const Cx bc = cmul({pa.beta_re, pa.beta_im}, {c.x, c.y});
                    y.re += bc.re; y.im += bc.im;
                    if (pa.gamma != 0.0) {
                        const double2 pv2 = __ldcg(reinterpret_cast<const double2*>(xp + ov));
                        y.re += pa.gamma * pv2.x; y.im += pa.gamma * pv2.y;
                    }
                }
                *reinterpret_cast<double2*>(xn + ov) = make_double2(y.re, y.im);
                Cx tt = y;
                if (pa.scale_term) tt = cmul({pa.c_re, pa.c_im}, y);
                const double2 so = __ldcg(reinterpret_cast<const double2*>(sum + ob));
                nw = {so.x + tt.re, so.y + tt.im};
                *reinterpret_cast<double2*>(sum + ob) = make_double2(nw.re, nw.im);
                mx = hypot(nw.re - so.x, nw.im - so.y);
            }
            const double ore = __shfl_xor_sync(0xffffffffu, nw.re, 1), oim = __shfl_xor_sync(0xffffffffu, nw.im, 1);
            const double omx = __shfl_xor_sync(0xffffffffu, mx, 1);
            if (side == 0) {
                mk = fmax(mk, mx); mb = fmax(mb, omx);
                dr += ore * nw.re + oim * nw.im;
                di += ore * nw.im - oim * nw.re;
            }
        }
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {       // lanes of equal particle (lane bit 1); bra lanes carry zeros
            mb = fmax(mb, __shfl_xor_sync(0xffffffffu, mb, off)); mk = fmax(mk, __shfl_xor_sync(0xffffffffu, mk, off));
            dr += __shfl_xor_sync(0xffffffffu, dr, off);          di += __shfl_xor_sync(0xffffffffu, di, off);
        }
        if (lane == 0 || lane == 2) {
            const int p = lane >> 1;
            wpart[w][p * 4 + 0] = mb; wpart[w][p * 4 + 1] = mk; wpart[w][p * 4 + 2] = dr; wpart[w][p * 4 + 3] = di;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            const int q = threadIdx.x;
            double v = wpart[0][q];
            for (int w2 = 1; w2 < TMA_THREADS / 32; ++w2) v = ((q & 3) < 2) ? fmax(v, wpart[w2][q]) : v + wpart[w2][q];
            S.blockpart[(size_t)b * 8 + q] = v;
        }

}

__global__ void __launch_bounds__(TMA_THREADS, 1)
series_kernel(const __grid_constant__ CUtensorMap tmap, const SeriesParams S)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar_full  = reinterpret_cast<uint64_t*>(smem + TmaSmem::off_bar_full);
    uint64_t* bar_empty = reinterpret_cast<uint64_t*>(smem + TmaSmem::off_bar_empty);
    double*   red       = reinterpret_cast<double*>(smem + TmaSmem::off_red);
    __shared__ Ctrl   sctrl;                           // every CTA keeps (and updates identically) its own copy
    __shared__ double wpart[TMA_THREADS / 32][8];
    __shared__ double fin[8];
    __shared__ PassParams spass;                       // this term's parameters (read by the epilogue and the decision)

    const MatvecParams& P = S.mv;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x, G = gridDim.x;
    const int t0 = int(((long long)P.T * b) / G), t1 = int(((long long)P.T * (b + 1)) / G);
    const int nt = t1 - t0;                            // host guarantees nt >= TMA_STAGES for every CTA
    const int panel0 = t0 / P.TPP, ct0 = t0 - panel0 * P.TPP;
    const uint64_t policy = policy_evict_first();
    const int r0 = int(((long long)P.M * b) / G), r1 = int(((long long)P.M * (b + 1)) / G);   // epilogue rows of this CTA

    if (threadIdx.x == 0) {
        prefetch_tensormap(&tmap);
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], N_CWARPS); }
        fence_barrier_init();
        sctrl = *S.ctrl;
    }
    __syncthreads();

    int prv = 2, cur = 0, nxt = 1;                     // vector rotation
    int s = 0, ph = 0, rslot = 0, sr = 0, phr = 0, rslotr = 0;   // ring state runs on across the terms
    unsigned long long bar_target = 0;

    // H' tiles of term 0 (the x_ket parts follow below, like for every later term)
    if (threadIdx.x == 0 && !sctrl.all_latched && S.n_steps > 0) {
        TileCursor tc = {panel0, ct0};
        for (int i = 0; i < TMA_STAGES; ++i) { issue_tile_h(smem, bar_full, &tmap, i, tc, policy); tc.next(P.TPP); }
    }

    for (int t = 0; t < S.n_steps; ++t) {
        if (sctrl.all_latched) break;                  // uniform over the grid: every CTA holds the same sctrl
        const bool more = (t + 1 < S.n_steps);
        if (threadIdx.x < sizeof(PassParams) / 8)        // visible to the CTA after the __syncthreads of barrier A
            reinterpret_cast<double*>(&spass)[threadIdx.x] = reinterpret_cast<const double*>(S.passes + t)[threadIdx.x];
        const double* xk_cur = pick3(S.vk, cur);
        const double* xb_cur = pick3(S.vb, cur);

        // x_ket of the tiles whose H' part was prefetched before the previous barriers (ring slots s, s+1, ...)
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy stores of other CTAs -> async-proxy reads
            TileCursor tc = {panel0, ct0};
            int slot = s;
            for (int i = 0; i < TMA_STAGES; ++i) { issue_tile_x(smem, bar_full, xk_cur, slot, tc); tc.next(P.TPP); if (++slot == TMA_STAGES) slot = 0; }
        }

        DYB_STAMP(0);
        // ---------------------------------------------------------------- dual product over this CTA's tiles
        ConsumerRegs r;
        zero_acc(r);
        int seg = P.seg_base[b];
        TileCursor cj = {panel0, ct0}, cr = {panel0, ct0}, cn = {panel0, ct0};
        for (int i = 0; i < TMA_STAGES; ++i) cn.next(P.TPP);           // tile jr + STAGES of this term
        TileCursor cw = {panel0, ct0};                                  // next term's tile (jr + STAGES - nt)

        for (int j = 0; j < nt + RETIRE_LAG; ++j) {
            if (j < nt) {
                if (j == 0 || cj.ct == 0) load_xb_coherent(r, xb_cur, cj.panel, w, lane);
                mbar_wait_bounded(&bar_full[s], uint32_t(ph));
                const uint8_t* stage = smem + (size_t)s * STAGE_BYTES;
                const double*  sH = reinterpret_cast<const double*>(stage);
                const double2* sX = reinterpret_cast<const double2*>(stage + STAGE_H_BYTES);
                double pv[TILE_COLS * NQ];
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
                const double tot = transpose_reduce<TILE_COLS * NQ>(pv, lane);
                if ((lane & RED_MASK) == 0)
                    red[rslot * (N_CWARPS * TILE_COLS * NQ) + w * (TILE_COLS * NQ) + (lane >> RED_SHIFT)] = tot;
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[s]);
                if (j == nt - 1 || cj.ct == P.TPP - 1) { store_acc(r, P.ket_slab, seg, w, lane); zero_acc(r); ++seg; }
                cj.next(P.TPP);
                if (++s == TMA_STAGES) { s = 0; ph ^= 1; }
                if (++rslot == RED_SLOTS) rslot = 0;
            }
            if (j >= RETIRE_LAG) {
                const int jr = j - RETIRE_LAG;
                if ((jr & (N_CWARPS - 1)) == w) {
                    mbar_wait_bounded(&bar_empty[sr], uint32_t(phr));
                    if (lane < TILE_COLS * NQ) {
                        const double* rs = red + rslotr * (N_CWARPS * TILE_COLS * NQ) + lane;
                        double sum = rs[0];
#pragma unroll
                        for (int ww = 1; ww < N_CWARPS; ++ww) sum += rs[ww * TILE_COLS * NQ];
                        P.bra_slab[((size_t)cr.panel * P.Ncpad + (size_t)cr.ct * TILE_COLS) * NQ + lane] = sum;
                    }
                    if (lane == 0) {
                        if (jr + TMA_STAGES < nt) { issue_tile_h(smem, bar_full, &tmap, sr, cn, policy); issue_tile_x(smem, bar_full, xk_cur, sr, cn); }
                        else if (more) issue_tile_h(smem, bar_full, &tmap, sr, cw, policy);     // next term: H' part only
                    }
                    __syncwarp();
                }
                if (jr + TMA_STAGES < nt) cn.next(P.TPP); else cw.next(P.TPP);
                cr.next(P.TPP);
                if (++sr == TMA_STAGES) { sr = 0; phr ^= 1; }
                if (++rslotr == RED_SLOTS) rslotr = 0;
            }
        }

        DYB_STAMP(1);
        bar_target += G;
        grid_barrier(S.gbar, bar_target);              // A: every slab of this term is complete

        DYB_STAMP(2);
        // ---------------------------------------------------------------- epilogue on rows r0..r1-1 of this CTA
        epilogue_phase(S, spass, sctrl, r0, r1, prv, cur, nxt, wpart);

        DYB_STAMP(3);
        bar_target += G;
        grid_barrier(S.gbar, bar_target);              // B: new vectors and every CTA's scalars are visible

        DYB_STAMP(4);
        // ---------------------------------------------------------------- replicated decision (same inputs, same order)
        {   // 256 threads: q = tid & 7, CTAs tid>>3, +32, ... (independent loads), then a fixed tree: same result in every CTA
            const int q = threadIdx.x & 7;
            const bool is_max = (q & 3) < 2;
            double v = 0.0;                              // maxima are of non-negative numbers: 0 is neutral for both
            for (int b0 = threadIdx.x >> 3; b0 < G; b0 += 128) {
                double x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = (b0 + 32 * u < G) ? __ldcg(S.blockpart + (size_t)(b0 + 32 * u) * 8 + q) : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) v = is_max ? fmax(v, x[u]) : v + x[u];
            }
#pragma unroll
            for (int off = 8; off < 32; off <<= 1) {
                const double o = __shfl_xor_sync(0xffffffffu, v, off);
                v = is_max ? fmax(v, o) : v + o;
            }
            if (lane < 8) wpart[w][lane] = v;
            __syncthreads();
            if (threadIdx.x < 8) {
                double f = wpart[0][q];
                for (int w2 = 1; w2 < TMA_THREADS / 32; ++w2) f = is_max ? fmax(f, wpart[w2][q]) : f + wpart[w2][q];
                fin[q] = f;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned dummy = sctrl.block_counter;
            apply_decision(&sctrl, spass, fin);
            sctrl.block_counter = dummy;
        }
        __syncthreads();
        const int old_prv = prv; prv = cur; cur = nxt; nxt = old_prv;
        DYB_STAMP(5);

        if (more && sctrl.all_latched) {
            // the series is decided but the next term's H' tiles are in flight: complete them (their barriers expect
            // the x_ket bytes too) before leaving, a CTA must not exit with copies landing in its shared memory
            if (threadIdx.x == 0) {
                const double* xk_next = pick3(S.vk, cur);
                asm volatile("fence.proxy.async;" ::: "memory");
                TileCursor tc = {panel0, ct0};
                int slot = s, par = ph;
                for (int i = 0; i < TMA_STAGES; ++i) {
                    issue_tile_x(smem, bar_full, xk_next, slot, tc); tc.next(P.TPP);
                    mbar_wait_bounded(&bar_full[slot], uint32_t(par));
                    if (++slot == TMA_STAGES) { slot = 0; par ^= 1; }
                }
            }
            __syncthreads();
        }
    }
    if (b == 0 && threadIdx.x == 0) { const unsigned keep = S.ctrl->block_counter; *S.ctrl = sctrl; S.ctrl->block_counter = keep; }
}

}  // namespace dyb
