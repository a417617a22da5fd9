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
    PassParams pass;
};

struct Cx { double re, im; };
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// fixed-order slab sums for owned row i, particle pp
__device__ __forceinline__ void reduce_slabs(const EpiParams& E, int i, int pp, Cx& hk, Cx& hb) {
    const int panel = i / PANEL_ROWS, il = i % PANEL_ROWS;
    const int s0 = E.pseg_start[panel], s1 = E.pseg_start[panel + 1];
    double kr = 0.0, ki = 0.0;
    for (int s = s0; s < s1; ++s) {
        const double2 v = *reinterpret_cast<const double2*>(E.ket_slab + (((size_t)s * PANEL_ROWS + il) * NQ + 2 * pp));
        kr += v.x; ki += v.y;
    }
    double br = 0.0, bi = 0.0;
    const size_t col = (size_t)E.row0 + i;
    for (int p = 0; p < E.n_bra_slabs; ++p) {
        const double2 v = *reinterpret_cast<const double2*>(E.bra_slab + (((size_t)p * E.Ncpad + col) * NQ + 2 * pp));
        br += v.x; bi += v.y;
    }
    hk = {kr, ki}; hb = {br, bi};
}

constexpr int EPI_THREADS = 256;

__global__ void __launch_bounds__(EPI_THREADS)
epilogue_kernel(const EpiParams E)
{
    __shared__ double wpart[EPI_THREADS / 32][8];
    __shared__ int    is_last;

    const int idx = blockIdx.x * EPI_THREADS + threadIdx.x;
    const int i = idx >> 1, pp = idx & 1;
    const PartPass pa = E.pass.part[pp];
    const bool live = (i < E.M) && pa.active && !E.ctrl->part[pp].latched;

    double mb = 0.0, mk = 0.0, dr = 0.0, di = 0.0;
    if (live) {
        Cx hk, hb;
        reduce_slabs(E, i, pp, hk, hb);
        const Cx alpha = {pa.alpha_re, pa.alpha_im};
        Cx yk = cmul(alpha, hk), yb = cmul(alpha, hb);
        const size_t ob = (size_t)i * NQ + 2 * pp;                       // owned-row slot (bra side, sums)
        const size_t ok = ((size_t)E.row0 + i) * NQ + 2 * pp;            // global-column slot (ket vectors)
        if (pa.three_term) {
            const Cx beta = {pa.beta_re, pa.beta_im};
            const double2 ck = *reinterpret_cast<const double2*>(E.cur_k + ok);
            const double2 cb = *reinterpret_cast<const double2*>(E.cur_b + ob);
            const double2 pk = *reinterpret_cast<const double2*>(E.prv_k + ok);
            const double2 pb = *reinterpret_cast<const double2*>(E.prv_b + ob);
            const Cx bk = cmul(beta, {ck.x, ck.y}), bb = cmul(beta, {cb.x, cb.y});
            yk.re += bk.re + pa.gamma * pk.x; yk.im += bk.im + pa.gamma * pk.y;
            yb.re += bb.re + pa.gamma * pb.x; yb.im += bb.im + pa.gamma * pb.y;
        }
        *reinterpret_cast<double2*>(E.nxt_k + ok) = make_double2(yk.re, yk.im);
        *reinterpret_cast<double2*>(E.nxt_b + ob) = make_double2(yb.re, yb.im);

        Cx tk = yk, tb = yb;
        if (pa.scale_term) { const Cx c = {pa.c_re, pa.c_im}; tk = cmul(c, yk); tb = cmul(c, yb); }
        const double2 sk = *reinterpret_cast<const double2*>(E.sum_k + ob);
        const double2 sb = *reinterpret_cast<const double2*>(E.sum_b + ob);
        const Cx nk = {sk.x + tk.re, sk.y + tk.im}, nb = {sb.x + tb.re, sb.y + tb.im};   // new = old + term
        *reinterpret_cast<double2*>(E.sum_k + ob) = make_double2(nk.re, nk.im);
        *reinterpret_cast<double2*>(E.sum_b + ob) = make_double2(nb.re, nb.im);
        mk = hypot(nk.re - sk.x, nk.im - sk.y);                          // abs(new - old), Taylor.f:300
        mb = hypot(nb.re - sb.x, nb.im - sb.y);
        dr = nb.re * nk.re + nb.im * nk.im;                              // conj(bra)*ket, dotc
        di = nb.re * nk.im - nb.im * nk.re;
    }

    // warp reduction that keeps the lane parity (= particle) separate; fmax ignores NaN like `abs(..)>tol`
#pragma unroll
    for (int off = 2; off < 32; off <<= 1) {
        mb = fmax(mb, __shfl_xor_sync(0xffffffffu, mb, off));
        mk = fmax(mk, __shfl_xor_sync(0xffffffffu, mk, off));
        dr += __shfl_xor_sync(0xffffffffu, dr, off);
        di += __shfl_xor_sync(0xffffffffu, di, off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 2) { wpart[warp][lane * 4 + 0] = mb; wpart[warp][lane * 4 + 1] = mk; wpart[warp][lane * 4 + 2] = dr; wpart[warp][lane * 4 + 3] = di; }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int t = threadIdx.x;
        double v = wpart[0][t];
        for (int w2 = 1; w2 < EPI_THREADS / 32; ++w2) v = ((t & 3) < 2) ? fmax(v, wpart[w2][t]) : v + wpart[w2][t];
        E.blockpart[(size_t)blockIdx.x * 8 + t] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&E.ctrl->block_counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;

    // ---- last block: final scalars in block order, then the decision the host used to take per term
    __threadfence();
    if (threadIdx.x < 8) {
        const int t = threadIdx.x;
        const volatile double* bp = E.blockpart;
        double v = bp[t];
        for (unsigned bb = 1; bb < gridDim.x; ++bb) v = ((t & 3) < 2) ? fmax(v, bp[(size_t)bb * 8 + t]) : v + bp[(size_t)bb * 8 + t];
        wpart[0][t] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        Ctrl* c = E.ctrl;
        for (int p = 0; p < 2; ++p) {
            const PartPass q = E.pass.part[p];
            PartState& st = c->part[p];
            if (!q.active || st.latched) continue;
            st.n_terms += 1;
            st.max_b = wpart[0][p * 4 + 0]; st.max_k = wpart[0][p * 4 + 1];
            st.dot_re = wpart[0][p * 4 + 2]; st.dot_im = wpart[0][p * 4 + 3];
            st.norm = hypot(st.dot_re, st.dot_im);
            const bool norm_ok = fabs(st.norm - q.norm_ref) < TOL_NORM;                    // Taylor.f:104,199
            if (q.check_conv) {
                const bool conv = !(st.max_b > TOL_TERM) && !(st.max_k > TOL_TERM);        // Taylor.f:194-195
                if (conv && norm_ok) { st.latched = 1; st.ok = 1; st.k_exit = q.k; }
                else if (q.last)     { st.latched = 1; st.ok = 0; st.k_exit = 0; }
            } else if (q.last) {
                st.latched = 1; st.ok = (q.last_ok_by_norm ? (norm_ok ? 1 : 0) : 1); st.k_exit = q.k;
            }
        }
        c->all_latched = (c->part[0].latched && c->part[1].latched) ? 1 : 0;
        c->block_counter = 0u;
        __threadfence();
    }
}

// ---- series start: optional adoption psi <- sum for particles whose last series succeeded, then
//      cur = psi, sum = psi for the particles that start a new series; resets the control block.
struct InitParams {
    int M, row0, Nc;
    int adopt[2], active[2];
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
            *reinterpret_cast<double2*>(I.cur_b + ob) = b; *reinterpret_cast<double2*>(I.sum_b + ob) = b;
            *reinterpret_cast<double2*>(I.cur_k + ok) = k; *reinterpret_cast<double2*>(I.sum_k + ob) = k;
        }
    }
    if (idx == 0) {
        for (int p = 0; p < 2; ++p) {
            PartState& st = I.ctrl->part[p];
            if (I.active[p]) { st.latched = 0; st.ok = 0; st.k_exit = 0; st.n_terms = 0; }
            else             { st.latched = 1; }
        }
        I.ctrl->all_latched = (I.active[0] || I.active[1]) ? 0 : 1;
        I.ctrl->block_counter = 0u;
    }
}

// ---- plain slab reduction into quad vectors (kernel-level parity entry dyb_dual_matvec)
__global__ void slab_reduce_kernel(const EpiParams E)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 1, pp = idx & 1;
    if (i >= E.M) return;
    Cx hk, hb;
    reduce_slabs(E, i, pp, hk, hb);
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
