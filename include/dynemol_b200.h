/* ============================================================================
 * dynemol_b200.h -- C ABI of the B200-native electron-hole wavepacket propagator
 *
 * Drop-in boundary for ONE hot path of lgcrego/Dynemol: the short-time
 * propagator of each nuclear step (form H' = S^-1 h, then the Taylor /
 * Chebyshev series of dense real-H' x complex-psi products for the electron and
 * hole wavepackets).  Plain pointers and sizes only; no torch / C++ types.
 *
 * Two groups of entry points:
 *
 *  (A) LEGACY FORTRAN SYMBOLS -- exactly the symbols the reference's Fortran
 *      binds through an implicit interface (ifort/ifx mangling: lower case,
 *      trailing underscore, every argument by reference, no hidden lengths).
 *      Each declaration cites the reference interface it replaces.
 *
 *  (B) NATIVE HANDLE API (dyb_*) -- the same path with device-resident state,
 *      used by the Python mirror (dynemol_b200/api.py), the tests and bench.py,
 *      and by a refreshed Fortran caller through iso_c_binding
 *      (INTEGRATION.md).  All dyb_* functions return 0 on success or a
 *      negative DYB_E* code; dyb_last_error() gives the message.
 *
 * Complex numbers are (re,im) pairs of doubles, i.e. Fortran complex*16 /
 * cuDoubleComplex layout.  Matrices are column-major (Fortran order).
 * The library never falls back to a CPU implementation: without a CUDA device
 * every compute entry fails (dyb_*: DYB_ENODEV; legacy: message + exit(1),
 * the reference's CHECK_INFO convention, GPU_Interface.cpp:58).
 * ========================================================================== */
#ifndef DYNEMOL_B200_H
#define DYNEMOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } dyb_complex;

/* ------------------------------------------------------------------ (A) legacy Fortran symbols */

/* Replaces Taylor_gpu.cpp:219-232 / 634-736 (identical prototype in
 * Chebyshev_gpu.cpp:232-245); called from ElHl_Chebyshev_GPU.f:269-272.
 * in : N; h_S, h_h host N x N col-major (lda = N); h_PSI_bra/ket N complex;
 *      *t_init, *t_max (ps); *tau (1/eV, read only).
 * out: h_H <- H' = S^-1 h (N x N); h_PSI_bra/ket advanced t_init -> t_max in
 *      place; h_AO_bra <- S^-1 PSI_bra (NOT conjugated, caller conjugates,
 *      ElHl_Chebyshev_GPU.f:304); h_AO_ket untouched; *save_tau <- converged
 *      tau of the first Convergence (Taylor.f:71).
 * Propagator semantics follow the CPU oracle Taylor.f:35-219. */
void propagationelhl_gpucaller_(const int* N, const double* h_S, const double* h_h, double* h_H,
                                dyb_complex* h_AO_bra, dyb_complex* h_AO_ket,
                                dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket,
                                const double* t_init, const double* t_max,
                                double* tau, double* save_tau);

/* NEW batched form of the above (SURVEY.md 8b "new symbol"): electron AND hole
 * in one call so that a single pass over H' serves both.  AO_* / PSI_* are
 * N x 2 col-major (column 1 electron, column 2 hole, ElHl_Chebyshev.f:228,253);
 * tau(2) read only, save_tau(2) written.  Replaces the two per-rank calls of
 * ElHl_Chebyshev_GPU.f:203-210,269-272. */
void propagationelhl2_gpucaller_(const int* N, const double* h_S, const double* h_h, double* h_H,
                                 dyb_complex* h_AO_bra, dyb_complex* h_AO_ket,
                                 dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket,
                                 const double* t_init, const double* t_max,
                                 double* tau, double* save_tau);

/* Replaces Taylor_gpu.cpp:295-330 (propagation only, H' given on the host). */
void propagation_gpucaller_(const int* n, double* tau, double* save_tau,
                            const double* t_init, const double* t_max,
                            dyb_complex* h_PSI_bra, dyb_complex* h_PSI_ket, const double* h_H);

/* Replaces Taylor_gpu.cpp:743-797 (called from diabatic-Ehren.f:115): K = X o A - H' A, all host N x N col-major. */
void ehrenfestkernel_gpu_(const int* N, const double* h_H, const double* h_A, const double* h_X, double* h_K);

/* Replaces Taylor_gpu.cpp:801-873 (exported by the reference, no Fortran caller): the same kernel from the AO packets
 * bra, ket (N x 2 complex, column 1 electron, column 2 hole); rho(i,j) = Re{ket(j,1) bra(i,1)} - Re{ket(j,2) bra(i,2)},
 * A = (rho + rho^T)/2 are formed on the device (calculate_rho / A_ad_nd of diabatic-Ehren.f:107-109). */
void ehrenfestkernel2_gpu_(const int* N, const dyb_complex* h_bra, const dyb_complex* h_ket, const double* h_H, const double* h_X, double* h_K);

/* Replaces Chebyshev_gpu.cpp:517-519. */
double nakedbessel_(const int* n, const double* x);

/* Replace GPU_Interface.cpp:129-134,226-302.  Exported WEAK so that a build
 * that still links the reference's GPU_Interface.o keeps its own definitions. */
void gpu_init_(const int* pid, const int* procs_per_dev);
void gpu_finalize_(void);
void gpu_pin_(void* ptr, int* size_bytes);
void gpu_unpin_(void* ptr);

/* Replace GPU_Interface.cpp:129-158 (xPU_syInvert :861-873, xPU_dsymm :574-627, xPU_dzgemv :405-496): the three
 * dispatchers Matrix_math.f routes the path's a2 / a3 / a7 through (Matrix_math.f:183-198, 79-121, 221-301).  WEAK.
 * Host matrices in and out (column-major); xpu_syinvert_ returns the full symmetric inverse (both triangles). */
void xpu_syinvert_(double* A, const char* UpLo, const int* N, int* info);
void xpu_dsymm_(const char* side, const char* UpLo, const int* M, const int* N, const double* alpha, double* hA, const int* LDA,
                double* hB, const int* LDB, const double* beta, double* hC, const int* LDC);
void xpu_dzgemv_(const char* transA, const int* M, const int* N, const dyb_complex* alpha, double* hA, const int* LDA,
                 dyb_complex* hX, const int* incX, const dyb_complex* beta, dyb_complex* hY, const int* incY);

/* ------------------------------------------------------------------ (B) native handle API */

#define DYB_OK        0
#define DYB_ENODEV   (-1)   /* no CUDA device / driver: there is no CPU fallback */
#define DYB_ECUDA    (-2)   /* CUDA runtime / cuBLAS / cuSOLVER / NCCL failure */
#define DYB_EINVAL   (-3)   /* bad argument or call order */
#define DYB_ENOMEM   (-4)
#define DYB_ESINGULAR (-5)  /* S is singular (LAPACK info > 0 in the reference) */

#define DYB_MODE_TAYLOR     0  /* reference-parity mode: Taylor.f:35-219 semantics */
#define DYB_MODE_CHEBYSHEV  1  /* Chebyshev/Bessel series on the spectrally rescaled H' */
/* Parity modes that reproduce the reference's GPU files decision for decision: one term fewer
 * per Taylor series, raw powers H'^k psi with c_k applied in the sum, cublasIdamax-style term test with a strict `<`
 * (Taylor_gpu.cpp:334-480,511-622), and the un-rescaled Chebyshev series of Chebyshev_gpu.cpp:347-485,524-643
 * (valid for tau * ||H'|| <~ 1 only, as in the reference).  Single GPU.  Checked on the B200 against the reference's
 * own GPU propagators compiled in place for sm_100a (tests/test_gpu_refgpu_modes.py). */
#define DYB_MODE_TAYLOR_REFGPU     2
#define DYB_MODE_CHEBYSHEV_REFGPU  3
/* ONE Chebyshev expansion per nuclear step on the rescaled H' (needs spectral bounds like DYB_MODE_CHEBYSHEV): the order
 * is taken from the decay of the Bessel coefficients J_k(dE*tau) (K ~ R + O(R^(1/3)) terms) instead of the reference's
 * cap of 25 (Chebyshev_gpu.cpp:119), with a single 1e-8 norm test at the end.  ~3.5x fewer passes over H' than the
 * order-25 chain at dt = 0.5 fs, and accurate to rounding instead of to the 1e-8 term test. */
#define DYB_MODE_CHEBYSHEV_FULL    4

#define DYB_KERNEL_AUTO 0
#define DYB_KERNEL_TMA  1   /* TMA + mbarrier staged persistent kernel (default) */
#define DYB_KERNEL_LDG  2   /* direct 128-bit global loads (baseline / cross-check) */

#define DYB_MAX_EVENTS 256

/* Decision trace of one particle for one propagate call (mirrors what the
 * oracle records, so parity tests can compare decisions, not only vectors). */
typedef struct {
    int32_t n_convergence_calls;
    int32_t n_substeps;
    int32_t n_matvec_pairs;      /* series terms this particle consumed */
    int32_t n_rescale;
    int32_t n_first_shrink;
    int32_t last_k_ref;
    int32_t n_events;
    int32_t ev_kind[DYB_MAX_EVENTS];   /* 1 = Convergence, 2 = steady sub-step */
    int32_t ev_k[DYB_MAX_EVENTS];
    int32_t ev_ok[DYB_MAX_EVENTS];
    double  ev_tau[DYB_MAX_EVENTS];
    double  norm_ref;
    double  final_tau;
} dyb_trace;

typedef struct dyb_ctx dyb_ctx;

const char* dyb_last_error(void);
const char* dyb_team_last_error(void);    /* message of the last failing dyb_team_* call on this thread */
int64_t dyb_legacy_passes_last(void);            /* el+hole terms of the last propagation made through a legacy symbol */
int64_t dyb_unwrap_pin_bytes(int size_bytes);   /* host-only: the byte count gpu_pin_ recovers from a wrapped Fortran default integer */
const char* dyb_version(void);
int  dyb_device_count(void);
/* Host-only (no device needed): the launch plan of the dual product.  out8 = {panels, tiles_per_panel, tiles, grid,
 * segments, tile_cols, panel_rows, padded_cols}; seg_base[grid] / pseg_start[panels+1] filled when non-NULL. */
int  dyb_plan(int N, int n_rows, int sm_count, int64_t* out8, int32_t* seg_base, int32_t* pseg_start);

/* Host-only: blocking of the shared-memory-resident series kernel (DYB_SERIES_RESIDENT) for an N x N operator.
 * out6 = {grid side, block size, smem column stride, dynamic smem bytes, threads per CTA, fits (0/1)}. */
int  dyb_resident_plan(int N, int sm_count, int64_t smem_optin_bytes, int64_t* out6);

/* Host-only: blocking of the streamed one-launch series kernel (DYB_SERIES_MID) for an N x N operator: a Gr x Gc grid of
 * CTAs, CTA (bi, bj) owns rows [bi*R, bi*R+R) x columns [bj*Cnp, bj*Cnp+Cnp) of H' and the E consecutive vector indices
 * starting at (bi*Gc + bj)*E.  out12 = {R, tile columns, Gr, Gc, Cnp, tiles per term, ring stages, E, owner CTAs, words an
 * owner collects per term (negative: its collect table holds 16-bit pairs), dynamic smem bytes, fits (0/1)}. */
int  dyb_mid_plan(int N, int sm_count, int64_t smem_optin_bytes, int64_t* out12);

/* Host-only: the tau of every remaining sub-step of the steady loop (Taylor.f:81-126: t += tau*h_bar, a last shorter
 * sub-step when less than one tau is left) assuming every norm test passes -- the schedule the library predicts when it
 * chains the sub-steps of a small operator into one launch.  Returns the number of sub-steps written (<= max_sub). */
int  dyb_steady_schedule(double t, double t_max, double tau, int max_sub, double* out_tau);
/* Host-only: the 25 series coefficients for a given tau and the number of terms the series would use.
 * DYB_MODE_TAYLOR: coefficient() of Taylor.f:224-239, k_max rule of :165-171.  DYB_MODE_CHEBYSHEV: Chebyshev_gpu.cpp:636-643
 * with R = de*tau and the phase of the spectral shift, k_max rule of :565-574. */
int  dyb_series_coefficients(int mode, double tau, double ebar, double de, dyb_complex* out25, int* k_max);

/* One context = one GPU, one basis size.  n_rows/row0 select a row shard of H'
 * (single GPU: row0 = 0, n_rows = N).  The context owns all device buffers. */
int  dyb_create(dyb_ctx** out, int device, int N, int row0, int n_rows);
int  dyb_destroy(dyb_ctx* ctx);
int  dyb_set_kernel(dyb_ctx* ctx, int kernel_variant);
/* How the terms of one series (one Convergence() call / one steady sub-step of Taylor.f:81-126) are launched:
 *   PER_TERM  two launches per term (dual product + fused epilogue, chained by programmatic dependent launch);
 *   RESIDENT  one cooperative launch per series, H' blocked over the shared memories of the SMs for the whole
 *             series (single GPU, N <= ~1800: the QM regions of the Ehrenfest / CSDM examples);
 *   MID       one cooperative launch per series, H' STREAMED per term by a TMA ring that runs across the terms, one grid
 *             barrier per term (single GPU, mid-size operators: resident range < N <~ 6000; csrc/mid.cuh).  Replaces the
 *             per-term GEMV launches and host round trips of Taylor_gpu.cpp:570-600 for the operators whose pass over H'
 *             (5 ... 40 us) is of the order of the launch overheads;
 *   AUTO      RESIDENT when the operator fits the shared memories, MID up to DYNEMOL_B200_MID_MAX (default 6144), else
 *             PER_TERM (default; env DYNEMOL_B200_SERIES=term|resident|mid|auto).
 * (Values 2 and 4 were the round-1 streaming-cooperative and streamed-block kernels: measured slower than PER_TERM
 * everywhere, removed; DESIGN.md keeps the measurements.) */
#define DYB_SERIES_AUTO     0
#define DYB_SERIES_PER_TERM 1
#define DYB_SERIES_RESIDENT 3
#define DYB_SERIES_MID      5
int  dyb_set_series_kernel(dyb_ctx* ctx, int kind);
int  dyb_get_info(dyb_ctx* ctx, int64_t* info16);   /* [0]=N [1]=ld [2]=n_rows [3]=grid [4]=tiles [5]=segments [6]=sm_count [7]=smem_bytes [8]=variant ... [13]=series kernel in effect [14]=resident grid side [15]=resident block size */

/* Operator: either upload H' (host, lda >= N; only rows row0..row0+n_rows-1 are kept),
 * or write it directly into the device buffer (bench: synthetic H' generated on the
 * device), or form it on the device from S and h (a2+a3 of SURVEY.md section 8). */
int  dyb_upload_hprime(dyb_ctx* ctx, const double* h_H, int64_t lda);
int  dyb_upload_hprime_device(dyb_ctx* ctx, const void* d_H, int64_t lda);   /* device -> device copy of rows row0.. */
/* device -> device copy of a block of owned rows: source holds ONLY rows local_row0..local_row0+n_rows-1
 * of the shard (n_rows x N, column-major, lda >= n_rows) */
int  dyb_upload_hprime_rows_device(dyb_ctx* ctx, const void* d_rows, int64_t lda, int local_row0, int n_rows);
int  dyb_hprime_device(dyb_ctx* ctx, void** d_ptr, int64_t* ld);
int  dyb_form_hprime(dyb_ctx* ctx, const double* h_S, const double* h_h, double* h_H_out /* may be NULL */);
int  dyb_form_hprime_device(dyb_ctx* ctx, const void* d_S, int64_t lds, const void* d_h, int64_t ldh);
/* Same, with the Hueckel matrix built on the device: h(i,j) = X_ij(IP, k_WH, V_shift) * S(i,j)
 * (Build_Huckel, ElHl_Chebyshev.f:296-323; X_ij, hamiltonians.f:33-63) so that only S crosses PCIe. */
int  dyb_form_hprime_from_overlap(dyb_ctx* ctx, const double* h_S, const double* IP, const double* k_WH,
                                  const double* V_shift, double* h_H_out /* may be NULL */);
int  dyb_download_hprime(dyb_ctx* ctx, double* h_H, int64_t lda);
/* device -> device copy OUT of the resident H': owned rows local_row0..local_row0+n_rows-1, all N columns, into a
 * column-major device buffer with leading dimension ldd >= n_rows (host layers use it to hand row blocks of an operator
 * formed on one GPU to the ranks of a row-sharded run). */
int  dyb_download_hprime_rows_device(dyb_ctx* ctx, void* d_dst, int64_t ldd, int local_row0, int n_rows);
/* dyb_form_hprime with the transfers overlapped (what the legacy symbols use): h goes up while S is being factorised, and
 * the download of H' into h_H_out starts when the solve ends and runs beside whatever is queued next (spectral bounds,
 * the series).  h_H_out is complete only after dyb_wait_outputs(). */
int  dyb_form_hprime_async(dyb_ctx* ctx, const double* h_S, const double* h_h, double* h_H_out /* may be NULL */);
int  dyb_wait_outputs(dyb_ctx* ctx);
/* Building blocks of the DISTRIBUTED formation used by dyb_team_form_hprime (N^3/3 + 2 N^3/P flops instead of 2.33 N^3 on
 * one GPU): the Cholesky factor of S on one full-size context; every row-sharded member solves S X = h for its block of
 * columns with a copy of the factor; the members then pull their row blocks out of the peers' column blocks. */
int  dyb_factor_overlap(dyb_ctx* ctx, const double* h_S);                        /* DYB_ESINGULAR if S is not SPD */
int  dyb_factor_device(dyb_ctx* ctx, void** d_U, int64_t* ldu);
int  dyb_upload_column_block(dyb_ctx* member, const double* h_h);                /* columns row0 .. row0+n_rows-1 of host h */
int  dyb_solve_column_block(dyb_ctx* member, const void* d_U, int64_t ldu, double* h_H_out /* host N x N or NULL */);
int  dyb_column_block_device(dyb_ctx* member, void** d_X);
int  dyb_take_rows_from_column_blocks(dyb_ctx* member, void* const* d_X, int n_blocks);

/* Wavepackets: n_part (1 or 2) columns of N complex, col-major. */
int  dyb_set_packets(dyb_ctx* ctx, int n_part, const dyb_complex* bra, const dyb_complex* ket);
int  dyb_get_packets(dyb_ctx* ctx, int n_part, dyb_complex* bra, dyb_complex* ket);

/* a4/a5: advance all particles from t_init to t_max.  tau[p] in (1/eV), save_tau[p] out.
 * traces may be NULL, else n_part entries. */
int  dyb_propagate(dyb_ctx* ctx, int mode, double t_init, double t_max,
                   const double* tau, double* save_tau, dyb_trace* traces);

/* Chebyshev mode (DYB_MODE_CHEBYSHEV): the reference's un-linked Chebyshev series (Chebyshev_gpu.cpp:347-485,
 * 524-643) on the spectrally rescaled operator (H' - ebar)/de; needs bounds [emin, emax] that enclose the
 * spectrum of H'.  Either pass them, or estimate them with n_iter Lanczos steps (each one pass of the dual
 * product) started from the current packets; `margin` widens the Ritz interval by that fraction on both sides. */
int  dyb_set_spectral_bounds(dyb_ctx* ctx, double emin, double emax);
int  dyb_get_spectral_bounds(dyb_ctx* ctx, double* emin, double* emax);
int  dyb_estimate_spectral_bounds(dyb_ctx* ctx, int n_iter, double margin, double* emin, double* emax);

/* Post-step quantities on the device (ElHl_Chebyshev.f:269-283):
 *   AO_bra = S^-1 Psi_bra (un-conjugated, as the legacy symbol returns it); needs dyb_form_hprime.
 *   populations: out[(n_frag+2) x n_part] = [t, frag pops..., total] with DUAL_bra = conj(ket),
 *   DUAL_ket = bra (data_output.f:87-147,242-263); fragment[i] in 0..n_frag-1 or -1. */
int  dyb_ao_bra(dyb_ctx* ctx, int n_part, dyb_complex* h_AO_bra);
int  dyb_populations(dyb_ctx* ctx, int n_part, int n_frag, const int32_t* fragment, double t, double* out);
/* QuasiParticleEnergies (ElHl_Chebyshev.f:329-371) = dotc(Psi_bra, H' Psi_ket) per particle; out = (re,im) pairs. */
int  dyb_quasiparticle_energies(dyb_ctx* ctx, int n_part, double* out_reim);
/* Diabatic-Ehrenfest kernel K = X o A - H' A (diabatic-Ehren.f:115-119) with the H' resident in the context;
 * A, X, K host N x N col-major. */
int  dyb_ehrenfest_kernel(dyb_ctx* ctx, const double* h_A, const double* h_X, double* h_K);
/* same, with A built on the device from the AO packets (N x 2 complex each): ehrenfestkernel2_gpu_ with H' resident */
int  dyb_ehrenfest_kernel2(dyb_ctx* ctx, const dyb_complex* h_bra, const dyb_complex* h_ket, const double* h_X, double* h_K);

/* Raw recursion for benchmarks and kernel-level parity: run n_terms el+hole series
 * terms (one pass over H' each, fused epilogue, no host decisions) starting from the
 * current packets, Taylor ratios of `tau`.  elapsed_ms (may be NULL) is measured with
 * CUDA events on the launching stream; kernel_ms (may be NULL) receives the matvec
 * kernel's own time summed over the terms (events around each launch). */
int  dyb_run_terms(dyb_ctx* ctx, double tau, int n_terms, float* elapsed_ms, float* kernel_ms);

/* One dual product with no epilogue state: ykt = H' xk (ket, 'N') and ybr = H'^T xb
 * (bra, 'T') for n_part columns; host in/out.  Kernel-level parity entry. */
int  dyb_dual_matvec(dyb_ctx* ctx, int n_part, const dyb_complex* xb, const dyb_complex* xk,
                     dyb_complex* yb, dyb_complex* yk);

/* Row-sharded H' over the GPUs of one box: one process per GPU, each holding rows
 * row0..row0+n_rows-1 (uniform: n_rows = N/world, row0 = rank*n_rows).  Rank 0 calls dyb_comm_unique_id, the host
 * layer ships the 128 bytes to every rank (torch.distributed / MPI), every rank calls dyb_comm_init.  Afterwards
 * dyb_set_packets / dyb_propagate / dyb_run_terms / dyb_get_packets are collective calls: per term the bra
 * partials are reduce-scattered and the new ket slices all-gathered with NCCL over NVLink. */
int  dyb_comm_unique_id(char* out128);
int  dyb_comm_init(dyb_ctx* ctx, int rank, int world, const char* id128);
/* Optional, after dyb_comm_init: fused exchange over NVLink peer memory.  Every rank exports one CUDA-IPC buffer
 * (64-byte handle), the host layer all-gathers the handles (rank order), every rank maps its peers.  From then on
 * each term's reduce-scatter (bra) and all-gather (ket) happen INSIDE the epilogue kernel as peer loads/stores,
 * with st.release.sys / ld.acquire.sys epoch flags instead of NCCL launches. */
int  dyb_comm_p2p_handle(dyb_ctx* ctx, char* out64);
int  dyb_comm_p2p_open(dyb_ctx* ctx, const char* handles);
int  dyb_comm_p2p_enable(dyb_ctx* ctx, int on);   /* collective: same value on every rank; 0 = NCCL collectives */
/* Peers that live in THIS process (dyb_team): phase 0 allocates the exchange buffer, phase 1 -- after every member did
 * phase 0 -- maps the peers by cudaDeviceEnablePeerAccess (no IPC).  members: `world` contexts in rank order. */
int  dyb_comm_p2p_open_local(dyb_ctx* ctx, dyb_ctx* const* members, int phase);

/* ------------------------------------------------------------------ single-process multi-GPU ("team")
 * the reference ABI is ONE host process, so a Fortran caller can only reach several GPUs if the library
 * drives them itself.  A team is `n_dev` row-sharded contexts (rank r on devices[r], rows r*N/n_dev ...) driven by one
 * host thread per device inside every call; H' = S^-1 h is formed on the first device and its row blocks are scattered
 * over NVLink; per term the fused peer-memory exchange of the row-sharded mode runs between the devices (peer access, no
 * IPC), NCCL (ncclCommInitRank per thread) serves the few collectives outside the term loop.  Results are those of the
 * one-process-per-GPU mode bit for bit.  The legacy symbols use a team when DYNEMOL_B200_GPUS=P > 1.
 * All calls take host buffers of the FULL problem (N x N, N x n_part), like the single-GPU API. */
typedef struct dyb_team dyb_team;
int  dyb_team_create(dyb_team** out, int n_dev, const int* devices /* NULL: 0..n_dev-1 */, int N);
int  dyb_team_destroy(dyb_team* team);
int  dyb_team_size(dyb_team* team);
int  dyb_team_form_hprime(dyb_team* team, const double* h_S, const double* h_h, double* h_H_out /* may be NULL; complete after dyb_team_wait_outputs */);
int  dyb_team_wait_outputs(dyb_team* team);
int  dyb_team_upload_hprime(dyb_team* team, const double* h_H, int64_t lda);
int  dyb_team_set_packets(dyb_team* team, int n_part, const dyb_complex* bra, const dyb_complex* ket);
int  dyb_team_get_packets(dyb_team* team, int n_part, dyb_complex* bra, dyb_complex* ket);
int  dyb_team_set_spectral_bounds(dyb_team* team, double emin, double emax);
int  dyb_team_estimate_spectral_bounds(dyb_team* team, int n_iter, double margin, double* emin, double* emax);
int  dyb_team_propagate(dyb_team* team, int mode, double t_init, double t_max, const double* tau, double* save_tau, dyb_trace* traces);
int  dyb_team_ao_bra(dyb_team* team, int n_part, dyb_complex* h_AO_bra);       /* needs dyb_team_form_hprime */
int  dyb_team_populations(dyb_team* team, int n_part, int n_frag, const int32_t* fragment, double t, double* out);
int  dyb_team_quasiparticle_energies(dyb_team* team, int n_part, double* out_reim);
int  dyb_team_run_terms(dyb_team* team, double tau, int n_terms, float* elapsed_ms_max);
int64_t dyb_team_passes_last(dyb_team* team);

int  dyb_sync(dyb_ctx* ctx);
int64_t dyb_launch_count(dyb_ctx* ctx);   /* kernels of THIS library launched so far */

#ifdef __cplusplus
}
#endif
#endif /* DYNEMOL_B200_H */
