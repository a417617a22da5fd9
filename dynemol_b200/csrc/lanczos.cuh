// lanczos.cuh -- device-side vector algebra of the spectral-bound estimate (dyb_estimate_spectral_bounds).
//
// The Chebyshev mode needs an interval [emin, emax] that encloses the spectrum of H' = S^-1 h.  It is estimated by a
// Lanczos run in the S inner product started from the wavepackets themselves (w0 = Psi_bra = S v0, v0 = Psi_ket):
// H'^T S = S H' keeps the left vectors w_j = S v_j, so the two-sided recurrence is the symmetric three-term one and
// every step costs ONE dual product of the hot kernel (H' v_j and H'^T w_j for electron and hole).  Round 1 kept the
// vectors on the host (one H2D + D2H of 4 N-vectors and O(j N) host loops per step); here everything but the final
// download of the tridiagonal coefficients stays on the device:
//
//   lz_dots_kernel     out[q] = <X_q | y>  for a list of vectors X_q (complex, both particles), fixed-order block sums
//   lz_combine_kernel  y = x - sum_q X_q c_q        (three-term step, Gram-Schmidt sweep, scaling)
//   lz_scalar_kernel   the few scalar decisions of a step (alpha_j, beta_{j+1}, breakdown test) taken by one thread
//
// Row-sharded operators: every vector is handled by its owned slice of rows; partial dot products are summed across the
// ranks by the caller (NCCL all-reduce) between lz_dots_kernel and its consumers.
// Full two-sided re-orthogonalisation (classical Gram-Schmidt, applied twice) against all previous vectors: without it
// the recurrence loses the duality w_j = S v_j after ~35 steps and produces Ritz values far outside the spectrum.
#pragma once
#include "common.cuh"

namespace dyb {

constexpr int LZ_THREADS = 512;
constexpr int LZ_MAX_IT  = 128;

// device-resident scalars of one run
struct LanczosState {
    double alpha[2][LZ_MAX_IT];        // diagonal of the tridiagonal matrix, per particle
    double beta[2][LZ_MAX_IT + 1];     // beta[p][j] couples steps j-1 and j (beta[p][0] = 0)
    int    ok[2][LZ_MAX_IT];           // step j has a sound successor (b2 above the breakdown threshold)
    int    alive[2];
    int    bad_start[2];               // <bra|ket> not positive: the packets are not an S-dual pair
};

// out[(q*2 + p)*2 + {0,1}] = sum_i conj(X_q[i,p]) * y[i,p]   (dotc: conjugated first argument), i over M owned rows.
// One block per q; strided per-thread sums and a fixed tree => bit-reproducible.
__global__ void __launch_bounds__(LZ_THREADS)
lz_dots_kernel(int M, const double* __restrict__ X, size_t x_stride, const double* __restrict__ y, double* __restrict__ out)
{
    __shared__ double sm[LZ_THREADS][4];
    const double* x = X + (size_t)blockIdx.x * x_stride;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < M; i += LZ_THREADS) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const double2 a = *reinterpret_cast<const double2*>(x + (size_t)i * NQ + 2 * p);
            const double2 b = *reinterpret_cast<const double2*>(y + (size_t)i * NQ + 2 * p);
            acc[2 * p]     += a.x * b.x + a.y * b.y;
            acc[2 * p + 1] += a.x * b.y - a.y * b.x;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sm[threadIdx.x][q] = acc[q];
    __syncthreads();
    for (int s = LZ_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s)
#pragma unroll
            for (int q = 0; q < 4; ++q) sm[threadIdx.x][q] += sm[threadIdx.x + s][q];
        __syncthreads();
    }
    if (threadIdx.x < 4) out[(size_t)blockIdx.x * 4 + threadIdx.x] = sm[0][threadIdx.x];
}

// y[i,p] = s_p * ( x[i,p] - sum_{q < nq} X_q[i,p] * c[q][p] )     (c complex, s real; x may alias y)
__global__ void lz_combine_kernel(int M, const double* x, const double* __restrict__ X, size_t x_stride, int nq,
                                  const double* __restrict__ coef /* [nq][2](re,im) */, const double* __restrict__ scale /* [2] or null */,
                                  double* y)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = idx >> 1, p = idx & 1;
    if (i >= M) return;
    const size_t o = (size_t)i * NQ + 2 * p;
    double2 v = *reinterpret_cast<const double2*>(x + o);
    for (int q = 0; q < nq; ++q) {
        const double2 a = *reinterpret_cast<const double2*>(X + (size_t)q * x_stride + o);
        const double cr = coef[(q * 2 + p) * 2], ci = coef[(q * 2 + p) * 2 + 1];
        v.x -= a.x * cr - a.y * ci;
        v.y -= a.x * ci + a.y * cr;
    }
    if (scale) { v.x *= scale[p]; v.y *= scale[p]; }
    *reinterpret_cast<double2*>(y + o) = v;
}

enum { LZ_OP_START = 0, LZ_OP_ALPHA = 1, LZ_OP_BETA = 2 };

// The scalar decisions of the run, one thread per particle (lanes are independent state machines):
//   START  d[p] = <w0|v0>: must be positive; scale[p] = 1/sqrt(d)
//   ALPHA  d[p] = <w_j|H'v_j>: alpha_j = Re d; coefficient list {alpha_j, beta_j} for the three-term step
//   BETA   d[p] = <w'|v'>: b2 = Re d; breakdown test b2 > 1e-24 (1 + alpha_j^2); beta_{j+1} = sqrt(b2); scale = 1/beta
// A dead particle gets zero scale factors: its vectors become zero and stay zero.
__global__ void lz_scalar_kernel(int op, int j, const double* __restrict__ d /* [2](re,im) */, LanczosState* st,
                                 double* __restrict__ coef_out, double* __restrict__ scale_out)
{
    const int p = threadIdx.x;
    if (p >= 2) return;
    const double re = d[2 * p];
    if (op == LZ_OP_START) {
        const bool good = re > 0.0;
        st->alive[p] = good ? 1 : 0; st->bad_start[p] = good ? 0 : 1;
        st->beta[p][0] = 0.0;
        scale_out[p] = good ? 1.0 / sqrt(re) : 0.0;
    } else if (op == LZ_OP_ALPHA) {
        st->alpha[p][j] = re;                         // list over {v_{j-1}, v_j} (contiguous in the vector store); j = 0: {v_0}
        const int qa = (j > 0) ? 1 : 0;
        if (j > 0) { coef_out[(0 * 2 + p) * 2] = st->beta[p][j]; coef_out[(0 * 2 + p) * 2 + 1] = 0.0; }
        coef_out[(qa * 2 + p) * 2] = re; coef_out[(qa * 2 + p) * 2 + 1] = 0.0;
    } else {
        const double a = st->alpha[p][j];
        const bool good = st->alive[p] && (re > 1e-24 * (1.0 + a * a));
        st->ok[p][j] = good ? 1 : 0;
        if (!good) st->alive[p] = 0;
        const double b = good ? sqrt(re) : 0.0;
        st->beta[p][j + 1] = b;
        scale_out[p] = good ? 1.0 / b : 0.0;
    }
}

}  // namespace dyb
