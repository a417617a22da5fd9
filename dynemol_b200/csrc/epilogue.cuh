// epilogue.cuh -- fused per-term epilogue of the series:
//   slab reduction (fixed order) -> recurrence update -> series accumulation -> max|term| and
//   <sum_bra|sum_ket> reductions (warp shuffles) -> on-device convergence / norm decision (last block).
//
// Replaces, per term, the reference's scal pre-pass, fused_Zxpby_and_subtract, 2x cublasIdamax+D2H,
// cublasZdotc (Taylor_gpu.cpp:570-600, Chebyshev_gpu_kernels.cu:296-313) and the host round-trips
// they imply; semantics follow the CPU oracle Taylor.f:182-211 / 90-104 (term = r*H'psi, new = old+term,
// isConverged on abs(new-old), norm test on abs(dotc(new_bra,new_ket))).
#pragma once
#include "common.cuh"

namespace dyb {

struct EpiParams {
    int M;                 // owned rows (single GPU: N)
    int row0;              // global index of the first owned row (bra slabs are indexed by global column)
    int n_bra_slabs;       // panels contributing to every bra entry
    int Ncpad;
    const double* ket_slab;
    const double* bra_slab;
    const int*    pseg_start;     // [n_panels+1] segment range of each panel
    const double* cur_b; const double* cur_k;   // x      (quad; cur_k indexed by global column)
    const double* prv_b; const double* prv_k;   // x_prev (three-term recurrences only)
    double* nxt_b; double* nxt_k;               // y
    double* sum_b; double* sum_k;               // running series sums (owned rows)
    double* blockpart;            // [grid][8] per-block partial scalars
    Ctrl*   ctrl;
    int     bra_col0;             // index of owned row 0 inside the bra slab rows (single GPU: row0; sharded: 0)
    int     defer_decision;       // sharded: write this rank's 8 scalars to scal_out, decide after the all-gather
    double* scal_out;
    PassParams pass;
};

struct Cx { double re, im; };
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// Fixed-order sum of `n` slab entries spaced `stride` doubles apart.  Loads are issued in batches of 8
// (independent, latency overlapped) and added strictly in slab order, so the result does not depend on
// the batching.
__device__ __forceinline__ Cx slab_sum(const double* __restrict__ base, size_t stride, int n) {
    double re = 0.0, im = 0.0;
    for (int s = 0; s < n; s += 8) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (s + u < n) v[u] = __ldcg(reinterpret_cast<const double2*>(base + (size_t)(s + u) * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (s + u < n) { re += v[u].x; im += v[u].y; }
    }
    return {re, im};
}

// H' x_ket at owned row i (sum over the panel's segments) / H'^T x_bra at global column row0+i (sum over panels)
__device__ __forceinline__ Cx reduce_ket(const EpiParams& E, int i, int pp) {
    const int panel = i / PANEL_ROWS, il = i % PANEL_ROWS;
    const int s0 = E.pseg_start[panel], s1 = E.pseg_start[panel + 1];
    return slab_sum(E.ket_slab + (((size_t)s0 * PANEL_ROWS + il) * NQ + 2 * pp), (size_t)PANEL_ROWS * NQ, s1 - s0);
}
__device__ __forceinline__ Cx reduce_bra(const EpiParams& E, int i, int pp) {
    return slab_sum(E.bra_slab + (((size_t)E.bra_col0 + i) * NQ + 2 * pp), (size_t)E.Ncpad * NQ, E.n_bra_slabs);
}

// The per-term decision the reference takes on the host (Taylor.f:194-207 inside Convergence, :102-105 in
// the steady loop), from the 8 reduced scalars v = {max_b, max_k, dot_re, dot_im} x {el, hl}.
__device__ __forceinline__ void decide_particle(PartState& st, const PartPass& q, const double* v4, bool allow_chain = true) {
    if (!q.active || st.latched) return;
    st.n_terms += 1;
    st.max_b = v4[0]; st.max_k = v4[1];
    st.dot_re = v4[2]; st.dot_im = v4[3];
    st.norm = hypot(st.dot_re, st.dot_im);
    const bool norm_ok = fabs(st.norm - q.norm_ref) < TOL_NORM;                        // Taylor.f:104,199
    if (q.check_conv) {
        // CPU variant: isConverged is false as soon as one modulus is > tol (Taylor.f:194-195,290-303); GPU variant: max_* hold
        // the modulus of the Idamax element and the test is a strict `<` (Taylor_gpu.cpp:581,587)
        const bool conv = q.test_gpu ? (st.max_b < TOL_TERM && st.max_k < TOL_TERM)
                                     : (!(st.max_b > TOL_TERM) && !(st.max_k > TOL_TERM));
        if (conv && norm_ok) { st.latched = 1; st.ok = 1; st.k_exit = q.k; }
        else if (q.last)     { st.latched = 1; st.ok = 0; st.k_exit = 0; }
    } else if (q.last) {
        const int ok = q.last_ok_by_norm ? (norm_ok ? 1 : 0) : 1;
        st.n_sub_ok += ok;
        if (!(ok && q.chain && allow_chain)) { st.latched = 1; st.ok = ok; st.k_exit = q.k; }   // a chained success just goes on
    }
}
__device__ __forceinline__ void apply_decision(Ctrl* c, const PassParams& pass, const double* v) {
    for (int p = 0; p < 2; ++p) decide_particle(c->part[p], pass.part[p], v + 4 * p);
    c->all_latched = (c->part[0].latched && c->part[1].latched) ? 1 : 0;
    c->block_counter = 0u;
    __threadfence();
}

constexpr int EPI_THREADS = 256;
// Reduced scalars of a term: slots p*4 + {0: max_b, 1: max_k, 2: dot_re, 3: dot_im} for particle p.  The GPU-variant term
// test (PartPass::test_gpu) needs an arg-max with a payload: slots 8 + 2p + {0: bra, 1: ket} carry the key
// max(|re|,|im|) of the winning element and slots p*4 + {0, 1} its modulus.  cublasIdamax returns the FIRST index of the
// maximum; here exact ties between different elements are resolved towards the larger modulus (no index is carried) --
// the two rules differ only when two elements tie bit for bit in their largest component AND straddle the tolerance.
constexpr int EPI_NS = 8, EPI_NS_RG = 12;
__device__ __forceinline__ void argmax_merge(double& key, double& val, double okey, double oval) {
    if (okey > key || (okey == key && oval > val)) { key = okey; val = oval; }
}
// a <- a (+) b over one slot set
template <int NS>
__device__ __forceinline__ void scal_merge(double* a, const double* b) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if constexpr (NS == EPI_NS_RG) {
            argmax_merge(a[8 + 2 * p], a[4 * p], b[8 + 2 * p], b[4 * p]);
            argmax_merge(a[9 + 2 * p], a[4 * p + 1], b[9 + 2 * p], b[4 * p + 1]);
        } else {
            a[4 * p] = fmax(a[4 * p], b[4 * p]); a[4 * p + 1] = fmax(a[4 * p + 1], b[4 * p + 1]);
        }
        a[4 * p + 2] += b[4 * p + 2]; a[4 * p + 3] += b[4 * p + 3];
    }
}
// threads per row = 4 (particle, side) x SL slab lanes; SL = 1 when few slabs contribute to a row (large N: ~20 segments
// per panel), SL = 4 when many do (small N: every CTA of the grid is a segment of the same panel)
constexpr int EPI_SL_WIDE = 4;

// =================================================================================================
// Row-sharded H', fused exchange + epilogue over NVLink peer memory (replaces ncclReduceScatter + epilogue +
// ncclAllGather + decide):
//   * bra:  y_bra[owned rows] = sum over ranks (fixed order) of the ranks' bra partial vectors, read straight from
//           the peers' HBM (P2P loads)                                                  == reduce-scatter
//   * ket:  the new ket slice is stored into EVERY rank's next ket vector (P2P stores)  == all-gather
//   * the 8 per-rank scalars are stored into every rank's table; after a flag handshake the last block of every
//           rank combines them in rank order and takes the (identical) decision.
// Flags are monotonically increasing 64-bit epochs written with st.release.sys into the peers' memory.
// =================================================================================================
constexpr int MAX_PEERS = 8;

struct PeerTable {
    int world, rank;
    const double* rs_send[MAX_PEERS];          // peers' bra partial vectors (this term's parity), full length
    double*       ket_next[MAX_PEERS];         // peers' next ket vector (global index)
    double*       scal_all[MAX_PEERS];         // peers' scalar tables (this term's parity): [world][8]
    unsigned long long* ready[MAX_PEERS];      // peers' "bra partials of rank r are ready" flags: slot [my rank]
    unsigned long long* done[MAX_PEERS];       // peers' "rank r finished its exchange" flags: slot [my rank]
    const unsigned long long* my_ready;        // my own flag arrays [world], written by the peers
    const unsigned long long* my_done;
    unsigned long long epoch;
};

__global__ void signal_ready_kernel(const PeerTable T)
{
    pdl_launch_dependents();
    pdl_wait();             // the local panel sum that precedes us has completed (programmatic dependent launch)
    // the bra partial vector of this rank is complete; publish it to every peer
    if (threadIdx.x < T.world) { __threadfence_system(); st_release_sys(T.ready[threadIdx.x], T.epoch); }
}

// strided slab sum: lane `sl` of SL adds slabs sl, sl+SL, ... (batches of 8 independent loads), fixed order
template <int SL>
__device__ __forceinline__ Cx slab_sum_strided(const double* __restrict__ base, size_t stride, int n, int sl) {
    double re = 0.0, im = 0.0;
    for (int s = sl; s < n; s += 8 * SL) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) if (s + SL * u < n) v[u] = __ldcg(reinterpret_cast<const double2*>(base + (size_t)(s + SL * u) * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) if (s + SL * u < n) { re += v[u].x; im += v[u].y; }
    }
    return {re, im};
}

// One kernel, two flavours.  P2P = false: single GPU, or the NCCL path of the row-sharded mode (bra partials arrive
// reduce-scattered in a one-slab buffer).  P2P = true: the exchange itself happens here (see the banner above).
// Thread layout: idx = ((row*2 + particle)*2 + side)*SL + slab-lane; side 0 = ket, 1 = bra.  The SL slab lanes split
// the slab (or peer) sum and merge it with shuffles; the two sides of a (row, particle) meet with one more shuffle
// for the <bra|ket> product.
template <bool P2P, int SL, bool RG = false>
__global__ void __launch_bounds__(EPI_THREADS)
epilogue_kernel_t(const EpiParams E, const PeerTable T)
{
    static_assert(!(P2P && RG), "the GPU-variant term test is single-GPU only");
    constexpr int NS = RG ? EPI_NS_RG : EPI_NS;
    __shared__ double wpart[EPI_THREADS / 32][NS];
    __shared__ int    is_last;

    if constexpr (P2P) {
        pdl_launch_dependents();        // the next dual product may be scheduled and prefetch its first H' tiles during the exchange
        pdl_wait();                     // signal_ready_kernel (and, transitively, the local panel sum) has completed
        __shared__ int s_abort;
        if (threadIdx.x == 0) s_abort = *reinterpret_cast<volatile int*>(&E.ctrl->peer_timeout);
        __syncthreads();
        if (s_abort) return;                                                             // an earlier term timed out: nothing is in step any more
        if (threadIdx.x < T.world && !wait_flag_ge(T.my_ready + threadIdx.x, T.epoch, &E.ctrl->peer_timeout)) s_abort = 1;   // all ranks' bra partials are visible
        __syncthreads();
        if (s_abort) return;
    } else {
        pdl_launch_dependents();        // the next dual product may start its H' prefetch while we run
        pdl_wait();                     // slabs of the dual product that precedes us are complete and visible
    }

    const int idx = blockIdx.x * EPI_THREADS + threadIdx.x;
    constexpr int LSL = (SL == 4) ? 2 : 0;
    static_assert(SL == 1 || SL == 4, "SL");
    const int i = idx >> (LSL + 2), pp = (idx >> (LSL + 1)) & 1, side = (idx >> LSL) & 1, sl = idx & (SL - 1);
    const PartPass pa = E.pass.part[pp];
    const bool live = (i < E.M) && pa.active && !E.ctrl->part[pp].latched;

    // ---- H' x (ket: this rank's segments; bra: panels / peers), split over the four slab lanes
    Cx hx = {0.0, 0.0};
    const size_t ob = (size_t)i * NQ + 2 * pp;                           // owned-row slot (bra vectors, both sums)
    const size_t og = ((size_t)E.row0 + i) * NQ + 2 * pp;                // global-index slot (ket vectors, peers' partials)
    if (live) {
        if (side == 0) {
            const int panel = i / PANEL_ROWS, il = i % PANEL_ROWS;
            const int s0 = E.pseg_start[panel], s1 = E.pseg_start[panel + 1];
            hx = slab_sum_strided<SL>(E.ket_slab + (((size_t)s0 * PANEL_ROWS + il) * NQ + 2 * pp), (size_t)PANEL_ROWS * NQ, s1 - s0, sl);
        } else if constexpr (P2P) {                                      // reduce-scatter by peer loads
            double2 v[MAX_PEERS];
#pragma unroll
            for (int r = 0; r < MAX_PEERS; ++r) if (r * SL + sl < T.world && r < MAX_PEERS / SL) v[r] = __ldcg(reinterpret_cast<const double2*>(T.rs_send[r * SL + sl] + og));
#pragma unroll
            for (int r = 0; r < MAX_PEERS; ++r) if (r * SL + sl < T.world && r < MAX_PEERS / SL) { hx.re += v[r].x; hx.im += v[r].y; }
        } else {
            hx = slab_sum_strided<SL>(E.bra_slab + (((size_t)E.bra_col0 + i) * NQ + 2 * pp), (size_t)E.Ncpad * NQ, E.n_bra_slabs, sl);
        }
    }
#pragma unroll
    for (int off = 1; off < SL; off <<= 1) {
        hx.re += __shfl_xor_sync(0xffffffffu, hx.re, off); hx.im += __shfl_xor_sync(0xffffffffu, hx.im, off);
    }

    // ---- recurrence, accumulation, term size: slab lane 0 of every (row, particle, side)
    double mx = 0.0;            // |new - old| of this thread's side
    double kx = 0.0;            // RG: max(|re|, |im|) of new - old (what cublasIdamax ranks, Taylor_gpu.cpp:84-89)
    Cx nw = {0.0, 0.0};         // new running sum of this thread's side
    if (live && sl == 0) {
        Cx y = cmul({pa.alpha_re, pa.alpha_im}, hx);
        const size_t ov = side ? ob : og;
        const double* cur = side ? E.cur_b : E.cur_k;
        const double* prv = side ? E.prv_b : E.prv_k;
        double* sum = side ? E.sum_b : E.sum_k;
        if (pa.three_term) {
            const double2 c = *reinterpret_cast<const double2*>(cur + ov);
            const Cx bc = cmul({pa.beta_re, pa.beta_im}, {c.x, c.y});
            y.re += bc.re; y.im += bc.im;
            if (pa.gamma != 0.0) {                     // first Chebyshev term has no x_prev (buffer may hold anything)
                const double2 pv = *reinterpret_cast<const double2*>(prv + ov);
                y.re += pa.gamma * pv.x; y.im += pa.gamma * pv.y;
            }
        }
        if (side) {
            *reinterpret_cast<double2*>(E.nxt_b + ob) = make_double2(y.re, y.im);
        } else if constexpr (P2P) {                                      // all-gather by peer stores
#pragma unroll
            for (int r = 0; r < MAX_PEERS; ++r)
                if (r < T.world) *reinterpret_cast<double2*>(T.ket_next[r] + og) = make_double2(y.re, y.im);
        } else {
            *reinterpret_cast<double2*>(E.nxt_k + og) = make_double2(y.re, y.im);
        }
        Cx t = y;
        if (pa.scale_term) t = cmul({pa.c_re, pa.c_im}, y);
        const double2 so = *reinterpret_cast<const double2*>(sum + ob);
        nw = {so.x + t.re, so.y + t.im};                                 // new = old + term   (Taylor.f:190-191)
        *reinterpret_cast<double2*>(sum + ob) = make_double2(nw.re, nw.im);
        mx = hypot(nw.re - so.x, nw.im - so.y);                          // abs(new - old)      (Taylor.f:300)
        if constexpr (RG) kx = fmax(fabs(nw.re - so.x), fabs(nw.im - so.y));
    }
    // partner lane = other side of the same (row, particle); idle lanes carry zeros
    const double ore = __shfl_xor_sync(0xffffffffu, nw.re, SL), oim = __shfl_xor_sync(0xffffffffu, nw.im, SL);
    const double omx = __shfl_xor_sync(0xffffffffu, mx, SL);
    double okx = 0.0;
    if constexpr (RG) okx = __shfl_xor_sync(0xffffffffu, kx, SL);
    double mb = 0.0, mk = 0.0, dr = 0.0, di = 0.0, kb = 0.0, kk = 0.0;
    if (side == 0 && sl == 0) {                                          // the ket lane owns the pair's contribution
        mk = mx; mb = omx; kk = kx; kb = okx;
        dr = ore * nw.re + oim * nw.im;                                  // conj(bra)*ket, dotc (Taylor.f:62,103,197)
        di = ore * nw.im - oim * nw.re;
    }
    // the rows of the warp (lane bits above the particle bit); particles (lane bit 2*SL) stay separate
#pragma unroll
    for (int off = 4 * SL; off < 32; off <<= 1) {
        if constexpr (RG) {
            const double ob = __shfl_xor_sync(0xffffffffu, mb, off), okb = __shfl_xor_sync(0xffffffffu, kb, off);
            const double ok_ = __shfl_xor_sync(0xffffffffu, mk, off), okk = __shfl_xor_sync(0xffffffffu, kk, off);
            argmax_merge(kb, mb, okb, ob); argmax_merge(kk, mk, okk, ok_);
        } else {
            mb = fmax(mb, __shfl_xor_sync(0xffffffffu, mb, off)); mk = fmax(mk, __shfl_xor_sync(0xffffffffu, mk, off));
        }
        dr += __shfl_xor_sync(0xffffffffu, dr, off);          di += __shfl_xor_sync(0xffffffffu, di, off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0 || lane == 2 * SL) {
        const int p = lane / (2 * SL);
        wpart[warp][p * 4 + 0] = mb; wpart[warp][p * 4 + 1] = mk; wpart[warp][p * 4 + 2] = dr; wpart[warp][p * 4 + 3] = di;
        if constexpr (RG) { wpart[warp][8 + 2 * p] = kb; wpart[warp][9 + 2 * p] = kk; }
    }
    __syncthreads();
    if constexpr (RG) {
        if (threadIdx.x == 0) {                                          // parity mode: one thread merges the 8 warps in order
            double v[NS];
#pragma unroll
            for (int t = 0; t < NS; ++t) v[t] = wpart[0][t];
            for (int w2 = 1; w2 < EPI_THREADS / 32; ++w2) scal_merge<NS>(v, wpart[w2]);
#pragma unroll
            for (int t = 0; t < NS; ++t) E.blockpart[(size_t)blockIdx.x * NS + t] = v[t];
        }
    } else if (threadIdx.x < 8) {
        const int t = threadIdx.x;
        double v = wpart[0][t];
        for (int w2 = 1; w2 < EPI_THREADS / 32; ++w2) v = ((t & 3) < 2) ? fmax(v, wpart[w2][t]) : v + wpart[w2][t];
        E.blockpart[(size_t)blockIdx.x * 8 + t] = v;
    }
    if constexpr (P2P) __threadfence_system();   // this block's peer stores (ket slices) are performed system-wide
    else __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&E.ctrl->block_counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;

    // ---- last block: final scalars (fixed strided order + fixed tree => deterministic), then the decision the
    //      host used to take per term
    __threadfence();
    __shared__ double fin[EPI_THREADS][NS];
    {
        double v[NS];
#pragma unroll
        for (int t = 0; t < NS; ++t) v[t] = 0.0;
        for (unsigned bb = threadIdx.x; bb < gridDim.x; bb += EPI_THREADS) {
            const double2* bp = reinterpret_cast<const double2*>(E.blockpart + (size_t)bb * NS);
            double a[NS];
#pragma unroll
            for (int t = 0; t < NS / 2; ++t) { const double2 q = __ldcg(bp + t); a[2 * t] = q.x; a[2 * t + 1] = q.y; }
            scal_merge<NS>(v, a);
        }
#pragma unroll
        for (int t = 0; t < NS; ++t) fin[threadIdx.x][t] = v[t];
        __syncthreads();
        for (int st = EPI_THREADS / 2; st > 0; st >>= 1) {
            if (threadIdx.x < st) scal_merge<NS>(fin[threadIdx.x], fin[threadIdx.x + st]);
            __syncthreads();
        }
    }
    if constexpr (P2P) {
        // my 8 scalars into slot [rank] of every peer's table, handshake, rank-ordered combination
        if (threadIdx.x < 8 * T.world) {
            const int r = threadIdx.x >> 3, t = threadIdx.x & 7;
            T.scal_all[r][(size_t)T.rank * 8 + t] = fin[0][t];
        }
        __threadfence_system();
        __syncthreads();
        __shared__ int s_abort2;
        if (threadIdx.x == 0) s_abort2 = 0;
        __syncthreads();
        if (threadIdx.x < T.world) {
            st_release_sys(T.done[threadIdx.x], T.epoch);                // "rank `rank` is done" on every peer
            if (!wait_flag_ge(T.my_done + threadIdx.x, T.epoch, &E.ctrl->peer_timeout)) s_abort2 = 1;   // and wait until every peer is done
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_abort2) { E.ctrl->block_counter = 0u; __threadfence(); }
        if (threadIdx.x == 0 && !s_abort2) {
            double v[8];
            const double* tab = T.scal_all[T.rank];
            for (int t = 0; t < 8; ++t) {
                double a = __ldcg(tab + t);
                for (int r = 1; r < T.world; ++r) { const double b = __ldcg(tab + (size_t)r * 8 + t); a = ((t & 3) < 2) ? fmax(a, b) : a + b; }
                v[t] = a;
            }
            apply_decision(E.ctrl, E.pass, v);
        }
    } else if (threadIdx.x == 0) {
        if (E.defer_decision) {                  // NCCL path of the row-sharded mode: decide after the all-gather
#pragma unroll
            for (int t = 0; t < 8; ++t) E.scal_out[t] = fin[0][t];
            E.ctrl->block_counter = 0u;
            __threadfence();
        } else {
            apply_decision(E.ctrl, E.pass, fin[0]);
        }
    }
}

// ---- series start: optional adoption psi <- sum for particles whose last series succeeded, then
//      cur = psi, sum = psi for the particles that start a new series; resets the control block.
struct InitParams {
    int M, row0, Nc;
    int adopt[2], active[2];
    int scale_sum[2];                 // Chebyshev: the running sum starts as c_0 * psi instead of psi
    double s_re[2], s_im[2];
    double* psi_b; double* psi_k;
    double* cur_b; double* cur_k;
    double* sum_b; double* sum_k;
    Ctrl* ctrl;
};

__global__ void series_init_kernel(const InitParams I)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 1, pp = idx & 1;
    if (i < I.M) {
        const size_t ob = (size_t)i * NQ + 2 * pp;
        const size_t ok = ((size_t)I.row0 + i) * NQ + 2 * pp;
        if (I.adopt[pp]) {
            *reinterpret_cast<double2*>(I.psi_b + ob) = *reinterpret_cast<const double2*>(I.sum_b + ob);
            *reinterpret_cast<double2*>(I.psi_k + ok) = *reinterpret_cast<const double2*>(I.sum_k + ob);
        }
        if (I.active[pp]) {
            const double2 b = *reinterpret_cast<const double2*>(I.psi_b + ob);
            const double2 k = *reinterpret_cast<const double2*>(I.psi_k + ok);
            *reinterpret_cast<double2*>(I.cur_b + ob) = b; *reinterpret_cast<double2*>(I.cur_k + ok) = k;
            double2 sb = b, sk = k;
            if (I.scale_sum[pp]) {
                const Cx c0 = {I.s_re[pp], I.s_im[pp]};
                const Cx tb = cmul(c0, {b.x, b.y}), tk = cmul(c0, {k.x, k.y});
                sb = make_double2(tb.re, tb.im); sk = make_double2(tk.re, tk.im);
            }
            *reinterpret_cast<double2*>(I.sum_b + ob) = sb; *reinterpret_cast<double2*>(I.sum_k + ob) = sk;
        }
    }
    if (idx == 0) {
        for (int p = 0; p < 2; ++p) {
            PartState& st = I.ctrl->part[p];
            if (I.active[p]) { st.latched = 0; st.ok = 0; st.k_exit = 0; st.n_terms = 0; st.n_sub_ok = 0; }
            else             { st.latched = 1; st.ok = 1; }      // sits this launch out (ok = 1: "did not fail")
        }
        I.ctrl->all_latched = (I.active[0] || I.active[1]) ? 0 : 1;
        I.ctrl->block_counter = 0u;
    }
}

// ---- row-sharded H': sum this rank's bra panels into one full-length partial vector (reduce-scatter input)
__global__ void bra_panel_reduce_kernel(int n_cols, int n_panels, int Ncpad, const double* __restrict__ bra_slab, double* __restrict__ out)
{
    pdl_launch_dependents();
    pdl_wait();             // no-op unless launched with programmatic stream serialization (fused peer-memory path)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one double2 (particle) of one column
    if (idx >= 2 * n_cols) return;
    const Cx v = slab_sum(bra_slab + (size_t)idx * 2, (size_t)Ncpad * NQ, n_panels);
    *reinterpret_cast<double2*>(out + (size_t)idx * 2) = make_double2(v.re, v.im);
}

// ---- row-sharded H': combine the per-rank scalars (rank order => identical on every rank) and decide
__global__ void decide_kernel(int world, const double* __restrict__ scal_all, Ctrl* ctrl, const PassParams pass)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double v[8];
    for (int t = 0; t < 8; ++t) {
        double a = scal_all[t];
        for (int r = 1; r < world; ++r) a = ((t & 3) < 2) ? fmax(a, scal_all[r * 8 + t]) : a + scal_all[r * 8 + t];
        v[t] = a;
    }
    apply_decision(ctrl, pass, v);
}

// ---- plain slab reduction into quad vectors (kernel-level parity entry dyb_dual_matvec)
__global__ void slab_reduce_kernel(const EpiParams E)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 1, pp = idx & 1;
    if (i >= E.M) return;
    const Cx hk = reduce_ket(E, i, pp), hb = reduce_bra(E, i, pp);
    *reinterpret_cast<double2*>(E.nxt_k + ((size_t)E.row0 + i) * NQ + 2 * pp) = make_double2(hk.re, hk.im);
    *reinterpret_cast<double2*>(E.nxt_b + (size_t)i * NQ + 2 * pp) = make_double2(hb.re, hb.im);
}

// ---- host complex columns (N x n_part, col-major)  <->  quad layout
__global__ void pack_quad_kernel(int N, int n_part, const double2* __restrict__ src, double* __restrict__ dst)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 1, pp = idx & 1;
    if (i >= N) return;
    const double2 v = (pp < n_part) ? src[(size_t)pp * N + i] : make_double2(0.0, 0.0);
    *reinterpret_cast<double2*>(dst + (size_t)i * NQ + 2 * pp) = v;
}
__global__ void unpack_quad_kernel(int N, int n_part, const double* __restrict__ src, double2* __restrict__ dst)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 1, pp = idx & 1;
    if (i >= N || pp >= n_part) return;
    dst[(size_t)pp * N + i] = *reinterpret_cast<const double2*>(src + (size_t)i * NQ + 2 * pp);
}

// ---- <bra|ket> per particle (dotc, Taylor.f:62): single block, fixed order => deterministic
__global__ void __launch_bounds__(1024)
dotc_kernel(int N, const double* __restrict__ b, const double* __restrict__ k, double* __restrict__ out4)
{
    __shared__ double sm[1024][4];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < N; i += 1024) {
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            const double2 vb = *reinterpret_cast<const double2*>(b + (size_t)i * NQ + 2 * pp);
            const double2 vk = *reinterpret_cast<const double2*>(k + (size_t)i * NQ + 2 * pp);
            acc[2 * pp]     += vb.x * vk.x + vb.y * vk.y;
            acc[2 * pp + 1] += vb.x * vk.y - vb.y * vk.x;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sm[threadIdx.x][q] = acc[q];
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s)
#pragma unroll
            for (int q = 0; q < 4; ++q) sm[threadIdx.x][q] += sm[threadIdx.x + s][q];
        __syncthreads();
    }
    if (threadIdx.x < 4) out4[threadIdx.x] = sm[0][threadIdx.x];
}

// ---- fragment populations, data_output.f:242-263 with DUAL_bra = conj(ket), DUAL_ket = bra
constexpr int MAX_FRAG = 30;
__global__ void __launch_bounds__(256)
populations_kernel(int N, int n_frag, const int* __restrict__ fragment, const double* __restrict__ psi_b,
                   const double* __restrict__ psi_k, double* __restrict__ out /* [2][MAX_FRAG+1] */)
{
    __shared__ double sm[256];
    const int pp = blockIdx.x;
    for (int f = 0; f <= n_frag; ++f) {          // f == n_frag: total
        double acc = 0.0;
        for (int i = threadIdx.x; i < N; i += 256) {
            if (f == n_frag || fragment[i] == f) {
                const double2 vb = *reinterpret_cast<const double2*>(psi_b + (size_t)i * NQ + 2 * pp);
                const double2 vk = *reinterpret_cast<const double2*>(psi_k + (size_t)i * NQ + 2 * pp);
                acc += vk.x * vb.x + vk.y * vb.y;          // Re( conj(ket) * bra )
            }
        }
        sm[threadIdx.x] = acc;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s]; __syncthreads(); }
        if (threadIdx.x == 0) out[pp * (MAX_FRAG + 1) + f] = sm[0];
        __syncthreads();
    }
}

}  // namespace dyb
