// propagator.cu -- host side of the B200-native electron-hole propagator: context, launch plan,
// the Taylor (reference-parity) and Chebyshev series drivers, H' = S^-1 h formation, and the C ABI
// declared in include/dynemol_b200.h.  Host logic mirrors, decision for decision, the reference's
// CPU oracle Taylor.f:35-219 as driven by ElHl_Chebyshev.f:174-276.
//
// There is deliberately no CPU fallback here: every compute entry needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include <cublas_v2.h>
#include <cusolverDn.h>
#include <nccl.h>       // types only: the library is bound with dlopen so that a process that already loaded
                        // torch's bundled libnccl.so.2 keeps using that one copy

#include "../../include/dynemol_b200.h"
#include "common.cuh"
#include "matvec.cuh"
#include "epilogue.cuh"
#include "resident.cuh"
#include "mid.cuh"
#include "lanczos.cuh"

using namespace dyb;
typedef std::complex<double> cplx;

static const int    ORDER = 25;           // Taylor.f:20
static const double H_BAR = 6.58264e-4;   // constants_m.f:23 (eV*ps)

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? DYB_ENODEV : DYB_ECUDA, \
                "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CKB(call) do { cublasStatus_t s_ = (call); if (s_ != CUBLAS_STATUS_SUCCESS) \
    return fail(DYB_ECUDA, "%s:%d %s: cublas status %d", __FILE__, __LINE__, #call, (int)s_); } while (0)
#define CKS(call) do { cusolverStatus_t s_ = (call); if (s_ != CUSOLVER_STATUS_SUCCESS) \
    return fail(DYB_ECUDA, "%s:%d %s: cusolver status %d", __FILE__, __LINE__, #call, (int)s_); } while (0)

// ------------------------------------------------------------------------------------------ NCCL (lazy)
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char*  (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.h) return DYB_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(DYB_ECUDA, "cannot load libnccl.so.2: %s", dlerror());
#define NSYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return fail(DYB_ECUDA, "libnccl lacks %s", name); } while (0)
    NSYM(GetUniqueId, "ncclGetUniqueId"); NSYM(CommInitRank, "ncclCommInitRank"); NSYM(CommDestroy, "ncclCommDestroy");
    NSYM(GetErrorString, "ncclGetErrorString"); NSYM(AllReduce, "ncclAllReduce"); NSYM(ReduceScatter, "ncclReduceScatter");
    NSYM(AllGather, "ncclAllGather"); NSYM(GroupStart, "ncclGroupStart"); NSYM(GroupEnd, "ncclGroupEnd");
#undef NSYM
    g_nccl.h = h;
    return DYB_OK;
}
#define CKN(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
    return fail(DYB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); } while (0)

// ------------------------------------------------------------------------------------------ context
struct dyb_ctx {
    int device = 0, sm_count = 0;
    int N = 0, row0 = 0, M = 0;          // basis size, first owned row, owned rows
    long long ld = 0;
    int NP = 0, TPP = 0, Ncpad = 0, grid = 0, n_seg = 0;
    int T = 0;
    size_t Lq = 0;                       // quad vector length (indices)
    int variant = DYB_KERNEL_TMA;
    bool use_pdl = true;                 // programmatic dependent launch between the dual product and the epilogue
    bool chain_steady = true;            // resident kernel: run the whole steady loop of a step in one launch (env DYNEMOL_B200_CHAIN=0 disables)
    int series_kind = DYB_SERIES_AUTO;   // how a series is launched: per term, streaming cooperative kernel, smem-resident kernel
    int res_Gd = 0, res_Bs = 0, res_ldS = 0;     // resident.cuh: grid side, block size, smem column stride (0: does not fit)
    size_t res_smem = 0;
    double *res_pk = nullptr, *res_pb = nullptr, *res_dscal = nullptr, *res_psi = nullptr;
    // mid.cuh: streamed one-launch series kernel for mid-size operators (plan + buffers; mid_fits: launchable on this device)
    bool mid_fits = false;
    int mid_auto_max = 6144;             // DYB_SERIES_AUTO selects it for resident range < N <= this (env DYNEMOL_B200_MID_MAX)
    double mid_l2_mb = 96.0;             // MB of H' the loads ask the L2 to keep (evict_last), env DYNEMOL_B200_MID_L2MB
    MidParams mid_P;                     // constant part of the launch parameters
    size_t mid_smem = 0;
    CUtensorMap tmap_mid;
    double *mid_pk = nullptr, *mid_pb = nullptr, *mid_xx = nullptr, *mid_sc = nullptr;   // epoch-tagged exchange buffers (16 B per double)
    size_t mid_bytes[4] = {0, 0, 0, 0};
    unsigned mid_epoch = 1;              // next unused epoch of the tagged words
    PassParams* d_passes = nullptr;      // per-term parameters of the series in flight
    unsigned long long* gbar = nullptr;  // grid barrier counter
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host <-> device copies that overlap work on `stream` (h upload behind potrf, H' download behind the series)
    cudaEvent_t ev_h_up = nullptr, ev_H_done = nullptr;
    std::thread out_thread;              // runs the (possibly host-blocking, pageable) download of H' beside the series
    int out_rc = 0;
    std::string out_err;
    CUtensorMap tmap;
    bool have_tmap = false;

    double* H = nullptr;                 // ld x N
    double* S = nullptr;                 // N x N factor of S (Cholesky or LU) kept for S^-1 applications
    double *ehr_A = nullptr, *ehr_X = nullptr, *ehr_K = nullptr, *ehr_vec = nullptr;   // Ehrenfest kernel scratch (3 N x N + packets), allocated on first use
    double* colblk = nullptr;            // N x M column block of h / H' (distributed formation on a team, dyb_solve_column_block)
    int64_t* ipiv = nullptr;             // LU pivots (fallback)
    bool have_factor = false, factor_is_lu = false;
    double *psi_b = nullptr, *psi_k = nullptr, *sum_b = nullptr, *sum_k = nullptr;
    double *vb[3] = {nullptr, nullptr, nullptr}, *vk[3] = {nullptr, nullptr, nullptr};
    double *ket_slab = nullptr, *bra_slab = nullptr, *blockpart = nullptr, *scal = nullptr, *io = nullptr;
    int *seg_base = nullptr, *pseg_start = nullptr, *frag = nullptr;
    Ctrl* ctrl = nullptr;                // device
    Ctrl* h_ctrl = nullptr;              // pinned host mirror
    double* h_scal = nullptr;            // pinned, 128 doubles
    int n_part = 0;
    bool have_bounds = false;
    double emin = 0.0, emax = 0.0;       // spectral bounds of H' for the Chebyshev mode
    double *lz_V = nullptr, *lz_W = nullptr, *lz_dots = nullptr;     // device Lanczos: vector stores [(n_iter+1)][M][NQ], dot/coefficient scratch
    LanczosState* lz_state = nullptr;
    int lz_cap = 0;                      // iterations the vector stores are sized for
    int64_t launches = 0;
    int64_t passes_last = 0;             // el+hole terms (passes over H') of the last propagate / run_terms
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t solver = nullptr;
    cusolverDnParams_t sparams = nullptr;
    std::vector<cudaEvent_t> ev;
    // row-sharded operation (one process per GPU): NCCL communicator + exchange buffers
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double *rs_send = nullptr, *rs_recv = nullptr, *scal_all = nullptr, *full_tmp = nullptr;
    // fused peer-memory exchange (NVLink P2P through CUDA IPC): one shared buffer per rank
    bool p2p = false;
    bool p2p_local = false;              // peers are contexts of this process (dyb_team): mapped by peer access, not by IPC
    char* comm_buf = nullptr;            // [rs_send x2 | ket vectors x3 | scalar tables x2 | ready flags | done flags]
    char* peer_base[MAX_PEERS] = {nullptr};
    size_t off_rs[2] = {0, 0}, off_vk[3] = {0, 0, 0}, off_scal[2] = {0, 0}, off_ready = 0, off_done = 0, comm_bytes = 0;
    unsigned long long epoch = 0;
    bool vk_in_comm = false;
};

static int ensure_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(DYB_ENODEV, "no CUDA device available (%s); dynemol_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(DYB_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    return DYB_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int build_tensor_map(dyb_ctx* c) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(DYB_ECUDA, "cuTensorMapEncodeTiled not available");
    // 3-D view of the column-major matrix: (row in sub-panel, sub-panel, column)
    cuuint64_t dims[3]    = {(cuuint64_t)SUB_ROWS, (cuuint64_t)(c->ld / SUB_ROWS), (cuuint64_t)c->N};
    cuuint64_t strides[2] = {(cuuint64_t)SUB_ROWS * 8, (cuuint64_t)c->ld * 8};
    cuuint32_t box[3]     = {(cuuint32_t)SUB_ROWS, (cuuint32_t)N_CWARPS, (cuuint32_t)TILE_COLS};
    cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = ((PFN_encodeTiled)fn)(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, c->H, dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DYB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    c->have_tmap = true;
    return DYB_OK;
}

// mid.cuh streams blocks of 256*WR rows x TC columns: same 3-D view of H', another box
static int build_tensor_map_mid(dyb_ctx* c, int WR, int TC) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(DYB_ECUDA, "cuTensorMapEncodeTiled not available");
    if (c->ld % MID_SUB) return fail(DYB_EINVAL, "leading dimension %lld is not a multiple of %d", c->ld, MID_SUB);
    cuuint64_t dims[3]    = {(cuuint64_t)MID_SUB, (cuuint64_t)(c->ld / MID_SUB), (cuuint64_t)c->N};
    cuuint64_t strides[2] = {(cuuint64_t)MID_SUB * 8, (cuuint64_t)c->ld * 8};
    cuuint32_t box[3]     = {(cuuint32_t)MID_SUB, (cuuint32_t)WR, (cuuint32_t)TC};
    cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = ((PFN_encodeTiled)fn)(&c->tmap_mid, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, c->H, dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DYB_ECUDA, "cuTensorMapEncodeTiled (mid) failed with CUresult %d", (int)r);
    return DYB_OK;
}

// Launch plan of the dual product (pure host arithmetic, also exported as dyb_plan for the CPU test-suite):
// tiles are dealt to CTAs as contiguous ranges of the panel-major tile list; a CTA range may straddle panel
// boundaries, each (CTA, panel) piece is a "segment" with its own ket slab; segments are globally ordered by panel.
struct Plan {
    int NP = 0, TPP = 0, Ncpad = 0, T = 0, grid = 0, n_seg = 0;
    std::vector<int> seg_base, pseg_start;
};
static Plan make_plan(int N, int M, int sm_count) {
    Plan p;
    p.NP    = (M + PANEL_ROWS - 1) / PANEL_ROWS;
    p.TPP   = (N + TILE_COLS - 1) / TILE_COLS;
    p.Ncpad = p.TPP * TILE_COLS;
    p.T     = p.NP * p.TPP;
    // small operators: fewer, longer-running CTAs (every CTA costs a 64 KiB ket slab, barrier set-up and a pipeline
    // fill, and every segment is one more term in the epilogue's slab sums)
    int min_tiles = (p.T < 8 * sm_count) ? 4 : 1;          // measured: N=1024 28 -> 21 us per term, N>=4096 unaffected
    if (const char* e = getenv("DYNEMOL_B200_MIN_TILES")) min_tiles = std::max(1, atoi(e));
    p.grid  = std::max(1, std::min(sm_count, p.T / min_tiles));
    p.seg_base.assign(p.grid, 0);
    std::vector<int> pcount(p.NP, 0);
    int seg = 0;
    for (int b = 0; b < p.grid; ++b) {
        const long long t0 = ((long long)p.T * b) / p.grid, t1 = ((long long)p.T * (b + 1)) / p.grid;
        p.seg_base[b] = seg;
        if (t1 <= t0) continue;
        const int p0 = (int)(t0 / p.TPP), p1 = (int)((t1 - 1) / p.TPP);
        for (int q = p0; q <= p1; ++q) { pcount[q]++; seg++; }
    }
    p.n_seg = seg;
    p.pseg_start.assign(p.NP + 1, 0);
    for (int q = 0; q < p.NP; ++q) p.pseg_start[q + 1] = p.pseg_start[q] + pcount[q];
    return p;
}

static int build_plan(dyb_ctx* c) {
    const Plan p = make_plan(c->N, c->M, c->sm_count);
    c->NP = p.NP; c->TPP = p.TPP; c->Ncpad = p.Ncpad; c->T = p.T; c->grid = p.grid; c->n_seg = p.n_seg;
    CK(cudaMalloc(&c->seg_base, sizeof(int) * c->grid));
    CK(cudaMalloc(&c->pseg_start, sizeof(int) * (c->NP + 1)));
    CK(cudaMemcpy(c->seg_base, p.seg_base.data(), sizeof(int) * c->grid, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->pseg_start, p.pseg_start.data(), sizeof(int) * (c->NP + 1), cudaMemcpyHostToDevice));
    return DYB_OK;
}

static int alloc_zero(double** p, size_t n_doubles) {
    CK(cudaMalloc(p, n_doubles * sizeof(double)));
    CK(cudaMemset(*p, 0, n_doubles * sizeof(double)));
    return DYB_OK;
}

// ------------------------------------------------------------------------------------------ launches
static MatvecParams matvec_params(dyb_ctx* c, const double* xk, const double* xb, bool use_ctrl) {
    MatvecParams P;
    P.M = c->M; P.Nc = c->N; P.ld = c->ld; P.TPP = c->TPP; P.T = c->T; P.Ncpad = c->Ncpad;
    P.H = c->H; P.Xk = xk; P.Xb = xb; P.ket_slab = c->ket_slab; P.bra_slab = c->bra_slab;
    P.seg_base = c->seg_base; P.ctrl = use_ctrl ? c->ctrl : nullptr;
    return P;
}

// Both hot kernels are launched with programmatic stream serialization (PDL): the successor may begin while the
// predecessor drains; each kernel executes griddepcontrol.wait before it touches the predecessor's output.
static cudaLaunchConfig_t pdl_config(dyb_ctx* c, unsigned grid, unsigned block, size_t smem, cudaLaunchAttribute* attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr->val.programmaticStreamSerializationAllowed = c->use_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cfg;
}

static int launch_matvec(dyb_ctx* c, const double* xk, const double* xb, bool use_ctrl) {
    const MatvecParams P = matvec_params(c, xk, xb, use_ctrl);
    cudaLaunchAttribute attr;
    if (c->variant == DYB_KERNEL_LDG) {
        cudaLaunchConfig_t cfg = pdl_config(c, c->grid, LDG_THREADS, 0, &attr);
        CK(cudaLaunchKernelEx(&cfg, dual_matvec_ldg_kernel, P));
    } else {
        cudaLaunchConfig_t cfg = pdl_config(c, c->grid, TMA_THREADS, TmaSmem::total, &attr);
        CK(cudaLaunchKernelEx(&cfg, dual_matvec_tma_kernel, c->tmap, P));
    }
    c->launches++;
    return DYB_OK;
}

static EpiParams epi_params(dyb_ctx* c, int cur, int prv, int nxt) {
    EpiParams E;
    E.M = c->M; E.row0 = c->row0; E.n_bra_slabs = c->NP; E.Ncpad = c->Ncpad;
    E.ket_slab = c->ket_slab; E.bra_slab = c->bra_slab; E.pseg_start = c->pseg_start;
    E.cur_b = c->vb[cur]; E.cur_k = c->vk[cur]; E.prv_b = c->vb[prv]; E.prv_k = c->vk[prv];
    E.nxt_b = c->vb[nxt]; E.nxt_k = c->vk[nxt]; E.sum_b = c->sum_b; E.sum_k = c->sum_k;
    E.blockpart = c->blockpart; E.ctrl = c->ctrl;
    E.bra_col0 = c->row0; E.defer_decision = 0; E.scal_out = nullptr;
    if (c->world > 1) {            // bra partials arrive reduce-scattered: one "slab" holding the owned slice
        E.bra_slab = c->rs_recv; E.n_bra_slabs = 1; E.bra_col0 = 0;
        E.defer_decision = 1; E.scal_out = c->scal_all + (size_t)c->rank * 8;
    }
    memset(&E.pass, 0, sizeof E.pass);
    return E;
}
// slab lanes of the epilogue: wide (4) when many segments contribute to every row (small N), else 1
static int epi_sl(const dyb_ctx* c) { return (c->n_seg > 32 * c->NP) ? EPI_SL_WIDE : 1; }
static int epi_grid(const dyb_ctx* c) { return (4 * epi_sl(c) * c->M + EPI_THREADS - 1) / EPI_THREADS; }

static int launch_epilogue(dyb_ctx* c, const EpiParams& E) {
    cudaLaunchAttribute attr;
    cudaLaunchConfig_t cfg = pdl_config(c, epi_grid(c), EPI_THREADS, 0, &attr);
    PeerTable none;
    memset(&none, 0, sizeof none);
    const bool rg = E.pass.part[0].test_gpu || E.pass.part[1].test_gpu;      // reference-GPU term test (parity modes)
    if (rg) {
        if (epi_sl(c) == 1) CK(cudaLaunchKernelEx(&cfg, epilogue_kernel_t<false, 1, true>, E, none));
        else CK(cudaLaunchKernelEx(&cfg, epilogue_kernel_t<false, EPI_SL_WIDE, true>, E, none));
    } else {
        if (epi_sl(c) == 1) CK(cudaLaunchKernelEx(&cfg, epilogue_kernel_t<false, 1, false>, E, none));
        else CK(cudaLaunchKernelEx(&cfg, epilogue_kernel_t<false, EPI_SL_WIDE, false>, E, none));
    }
    c->launches++;
    return DYB_OK;
}

// One el+hole series term.  Single GPU: dual product + fused epilogue.  Row-sharded: dual product on the local
// rows, reduce-scatter of the bra partials, epilogue on the owned slice, all-gather of the new ket slice and of
// the per-rank scalars, replicated decision.
static int run_term(dyb_ctx* c, const EpiParams& E, int cur, int nxt, bool use_ctrl) {
    int rc;
    if ((rc = launch_matvec(c, c->vk[cur], c->vb[cur], use_ctrl))) return rc;
    if (c->world == 1) return launch_epilogue(c, E);
    const int n2 = 2 * c->N;
    if (c->p2p) {
        // fused exchange over NVLink peer memory: local panel sum -> publish -> one kernel does reduce-scatter (peer
        // loads), epilogue, all-gather (peer stores), scalar exchange and the replicated decision.  All four launches of a
        // term are chained by programmatic dependent launch: each kernel is scheduled while its predecessor drains and
        // blocks in griddepcontrol.wait until that one has completed, so the launch latencies (and, for the next dual
        // product, the TMA prefetch of its first H' tiles) overlap the exchange.
        const unsigned long long epoch = ++c->epoch;
        const int par = (int)(epoch & 1);
        double* my_rs = reinterpret_cast<double*>(c->comm_buf + c->off_rs[par]);
        cudaLaunchAttribute attr;
        {
            cudaLaunchConfig_t cfg = pdl_config(c, (n2 + 255) / 256, 256, 0, &attr);
            CK(cudaLaunchKernelEx(&cfg, bra_panel_reduce_kernel, c->N, c->NP, c->Ncpad, (const double*)c->bra_slab, my_rs));
            c->launches++;
        }
        PeerTable T;
        memset(&T, 0, sizeof T);
        T.world = c->world; T.rank = c->rank; T.epoch = epoch;
        for (int r = 0; r < c->world; ++r) {
            char* pb = c->peer_base[r];
            T.rs_send[r]  = reinterpret_cast<const double*>(pb + c->off_rs[par]);
            T.ket_next[r] = reinterpret_cast<double*>(pb + c->off_vk[nxt]);
            T.scal_all[r] = reinterpret_cast<double*>(pb + c->off_scal[par]);
            T.ready[r]    = reinterpret_cast<unsigned long long*>(pb + c->off_ready) + c->rank;
            T.done[r]     = reinterpret_cast<unsigned long long*>(pb + c->off_done) + c->rank;
        }
        T.my_ready = reinterpret_cast<const unsigned long long*>(c->comm_buf + c->off_ready);
        T.my_done  = reinterpret_cast<const unsigned long long*>(c->comm_buf + c->off_done);
        {
            cudaLaunchConfig_t cfg = pdl_config(c, 1, 32, 0, &attr);
            CK(cudaLaunchKernelEx(&cfg, signal_ready_kernel, T));
            c->launches++;
        }
        EpiParams E2 = E;
        E2.defer_decision = 0;
        {
            cudaLaunchConfig_t cfg = pdl_config(c, epi_grid(c), EPI_THREADS, 0, &attr);
            if (epi_sl(c) == 1) CK(cudaLaunchKernelEx(&cfg, epilogue_kernel_t<true, 1, false>, E2, T));
            else CK(cudaLaunchKernelEx(&cfg, epilogue_kernel_t<true, EPI_SL_WIDE, false>, E2, T));
            c->launches++;
        }
        return DYB_OK;
    }
    bra_panel_reduce_kernel<<<(n2 + 255) / 256, 256, 0, c->stream>>>(c->N, c->NP, c->Ncpad, c->bra_slab, c->rs_send);
    c->launches++;
    CK(cudaGetLastError());
    CKN(g_nccl.ReduceScatter(c->rs_send, c->rs_recv, (size_t)c->M * NQ, ncclDouble, ncclSum, c->comm, c->stream));
    if ((rc = launch_epilogue(c, E))) return rc;
    CKN(g_nccl.GroupStart());
    CKN(g_nccl.AllGather(c->vk[nxt] + (size_t)c->row0 * NQ, c->vk[nxt], (size_t)c->M * NQ, ncclDouble, c->comm, c->stream));
    CKN(g_nccl.AllGather(c->scal_all + (size_t)c->rank * 8, c->scal_all, 8, ncclDouble, c->comm, c->stream));
    CKN(g_nccl.GroupEnd());
    decide_kernel<<<1, 32, 0, c->stream>>>(c->world, c->scal_all, c->ctrl, E.pass);
    c->launches++;
    CK(cudaGetLastError());
    return DYB_OK;
}

// Operator resident in shared memory for the whole series (resident.cuh): small N only, single GPU.
static bool resident_ok(const dyb_ctx* c) {
    return (c->series_kind == DYB_SERIES_RESIDENT || c->series_kind == DYB_SERIES_AUTO) && c->world == 1 && c->res_Gd > 0;
}
constexpr int MAX_SERIES_TERMS = 32;
constexpr int MAX_CHAIN_PASSES = 4096;      // chained steady sub-steps of one launch (resident kernel)

static int run_series_resident(dyb_ctx* c, const std::vector<PassParams>& passes) {
    const int n = (int)passes.size();
    if (n < 1 || n > MAX_CHAIN_PASSES) return fail(DYB_EINVAL, "series length %d out of range", n);
    CK(cudaMemcpyAsync(c->d_passes, passes.data(), sizeof(PassParams) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->gbar, 0, sizeof(unsigned long long), c->stream));
    ResidentParams R;
    memset(&R, 0, sizeof R);
    R.H = c->H; R.ld = c->ld; R.N = c->N; R.Gd = c->res_Gd; R.Bs = c->res_Bs; R.ldS = c->res_ldS;
    R.x0k = c->vk[0]; R.x0b = c->vb[0]; R.sum_b = c->sum_b; R.sum_k = c->sum_k;
    R.pk = c->res_pk; R.pb = c->res_pb; R.dscal = c->res_dscal; R.psi_store = reinterpret_cast<double2*>(c->res_psi);
    R.ctrl = c->ctrl; R.passes = c->d_passes; R.n_steps = n; R.gbar = c->gbar;
    bool rg = false;                           // reference-GPU term test (parity modes): the instantiation that carries arg-max keys
    for (const PassParams& pp : passes) if (pp.part[0].test_gpu || pp.part[1].test_gpu) { rg = true; break; }
#ifdef DYB_SERIES_PROF
    static long long* d_rprof = nullptr;
    const int rgrid = c->res_Gd * c->res_Gd;
    const size_t n_rprof = (size_t)MAX_SERIES_TERMS * rgrid * 6;
    if (!d_rprof) CK(cudaMalloc(&d_rprof, (size_t)MAX_SERIES_TERMS * 512 * 6 * 8));
    CK(cudaMemsetAsync(d_rprof, 0, n_rprof * 8, c->stream));
    R.prof = d_rprof;
#endif
    void* args[] = {(void*)&R};
    CK(cudaLaunchCooperativeKernel(rg ? (const void*)resident_series_kernel_t<true> : (const void*)resident_series_kernel_t<false>,
                                   dim3(c->res_Gd * c->res_Gd), dim3(RES_THREADS), args, c->res_smem, c->stream));
    c->launches++;
#ifdef DYB_SERIES_PROF
    {   // diagnostic build: mean / max cycles of each phase over CTAs and terms
        static int calls = 0;
        if (calls++ % 50 == 1) {
            std::vector<long long> h(n_rprof);
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaMemcpy(h.data(), d_rprof, n_rprof * 8, cudaMemcpyDeviceToHost));
            const char* name[6] = {"product", "partial-store", "barrier", "gather+decide", "update+diag", "loop-gap"};
            double mean[6] = {0}, mx[6] = {0};
            for (int t = 0; t + 1 < std::min(n, MAX_SERIES_TERMS); ++t) for (int b = 0; b < rgrid; ++b) {
                const long long* q = &h[((size_t)t * rgrid + b) * 6];
                for (int i = 0; i < 6; ++i) {
                    const long long nx = (i < 5) ? q[i + 1] : h[((size_t)(t + 1) * rgrid + b) * 6];
                    const double d = double(nx - q[i]);
                    mean[i] += d; mx[i] = std::max(mx[i], d);
                }
            }
            fprintf(stderr, "resident_prof N=%d grid=%d Bs=%d terms=%d:", c->N, rgrid, c->res_Bs, n);
            for (int i = 0; i < 6; ++i) fprintf(stderr, "  %s mean %.0f max %.0f cyc;", name[i], mean[i] / ((double)(n - 1) * rgrid), mx[i]);
            fprintf(stderr, "\n");
        }
    }
#endif
    return DYB_OK;
}

// Mid-size operators streamed by ONE launch per series (mid.cuh): single GPU; not for the reference-GPU term test.
// Blocking (pure host arithmetic, exported as dyb_mid_plan for the CPU tests): R = 512 rows per CTA, Gr = ceil(N / R) block
// rows, Gc = min(64, sm_count / Gr) block columns of Cnp = roundup(ceil(N / Gc), 8) columns, Gc = ceil(N / Cnp); every CTA
// owns E = ceil(N / (Gr Gc)) consecutive indices of the vectors (n_own = ceil(N / E) CTAs own at least one).
struct MidPlan { int WR, R, TC, Gr, Gc, Cnp, NT, ST, E, n_own, tab16; size_t smem; bool fits; };
static MidPlan make_mid_plan(int N, int sm_count, size_t smem_optin, size_t static_smem) {
    MidPlan m;
    memset(&m, 0, sizeof m);
    m.WR = 2; m.R = m.WR * MID_SUB; m.TC = (MID_WARPS / m.WR) * MID_CPW;
    m.Gr = (N + m.R - 1) / m.R;
    const int gc0 = std::min(64, sm_count / std::max(1, m.Gr));        // <= 64 partials per ket entry (summed by one thread)
    if (gc0 < 1) return m;
    const int cn = (N + gc0 - 1) / gc0;
    m.Cnp = (cn + m.TC - 1) / m.TC * m.TC;
    m.Gc = (N + m.Cnp - 1) / m.Cnp;
    m.NT = m.Cnp / m.TC;
    const int G = m.Gr * m.Gc;
    m.E = (N + G - 1) / G;
    m.n_own = (N + m.E - 1) / m.E;
    const size_t budget = std::min((size_t)MID_SMEM_MAX, smem_optin > static_smem ? smem_optin - static_smem : 0);
    // the collect table with 32-bit offsets is the faster one; the 16-bit form takes over when it saves a ring stage below 5
    int st32 = MID_MAX_ST, st16 = MID_MAX_ST;
    while (st32 >= 2 && (size_t)MidSmem(st32, m.R, m.Cnp, m.E, m.Gr + m.Gc, m.n_own, 0).total > budget) --st32;
    while (st16 >= 2 && (size_t)MidSmem(st16, m.R, m.Cnp, m.E, m.Gr + m.Gc, m.n_own, 1).total > budget) --st16;
    m.tab16 = (st16 > st32 && st32 < 4) ? 1 : 0;
    m.ST = m.tab16 ? st16 : st32;
    m.smem = (size_t)MidSmem(std::max(m.ST, 1), m.R, m.Cnp, m.E, m.Gr + m.Gc, m.n_own, m.tab16).total;
    // an owner's indices span at most two block columns (E <= Cnp); the 8 scalar slots of <= 16 * MID_SCU owners; a consumer's
    // R + Cnp entries in MID_XU words per thread; the bra partials of the WR row groups in the union region
    m.fits = m.ST >= 2 && G <= sm_count && m.E <= MID_MAX_E && m.E <= m.Cnp && m.n_own <= 16 * MID_SCU
             && (m.Cnp + m.R) * NQ <= MID_XT * MID_XU && 2 * m.E * NQ <= 3 * MID_XT
             && (m.Gr + m.Gc) * m.E * NQ <= MID_THREADS * MID_CWP && m.n_own * 8 <= MID_THREADS * MID_CWS && (size_t)m.WR * m.Cnp * NQ * 8 <= (size_t)MID_U_BYTES;
    return m;
}
static bool mid_ok(const dyb_ctx* c, bool refgpu) {
    if (!c->mid_fits || c->world != 1 || refgpu) return false;
    if (c->series_kind == DYB_SERIES_MID) return true;
    return c->series_kind == DYB_SERIES_AUTO && c->res_Gd == 0 && c->N <= c->mid_auto_max;
}

static int run_series_mid(dyb_ctx* c, const std::vector<PassParams>& passes) {
    const int n = (int)passes.size();
    if (n < 1 || n > MAX_CHAIN_PASSES) return fail(DYB_EINVAL, "series length %d out of range", n);
    CK(cudaMemcpyAsync(c->d_passes, passes.data(), sizeof(PassParams) * n, cudaMemcpyHostToDevice, c->stream));
    MidParams P = c->mid_P;
    if (c->mid_epoch > 0xfff00000u) {          // the 32-bit epochs are about to wrap: forget every tagged word (stream-ordered)
        CK(cudaMemsetAsync(c->mid_pk, 0, c->mid_bytes[0], c->stream)); CK(cudaMemsetAsync(c->mid_pb, 0, c->mid_bytes[1], c->stream));
        CK(cudaMemsetAsync(c->mid_xx, 0, c->mid_bytes[2], c->stream)); CK(cudaMemsetAsync(c->mid_sc, 0, c->mid_bytes[3], c->stream));
        c->mid_epoch = 1;
    }
    P.epoch0 = c->mid_epoch;
    c->mid_epoch += (unsigned)n + 4u;
    P.x0k = c->vk[0]; P.x0b = c->vb[0]; P.sum_b = c->sum_b; P.sum_k = c->sum_k;
    P.pk = reinterpret_cast<ulonglong2*>(c->mid_pk); P.pb = reinterpret_cast<ulonglong2*>(c->mid_pb);
    P.xx = reinterpret_cast<ulonglong2*>(c->mid_xx); P.sc = reinterpret_cast<ulonglong2*>(c->mid_sc);
    P.ctrl = c->ctrl; P.passes = c->d_passes; P.n_steps = n;
#ifdef DYB_SERIES_PROF
    static long long* d_mprof = nullptr;
    const int mgrid = P.Gr * P.Gc;
    const size_t n_mprof = (size_t)MAX_SERIES_TERMS * mgrid * 16;
    if (!d_mprof) CK(cudaMalloc(&d_mprof, (size_t)MAX_SERIES_TERMS * 512 * 16 * 8));
    CK(cudaMemsetAsync(d_mprof, 0, n_mprof * 8, c->stream));
    P.prof = d_mprof;
#else
    P.prof = nullptr;
#endif
    void* args[] = {(void*)&c->tmap_mid, (void*)&P};
    CK(cudaLaunchCooperativeKernel((const void*)mid_series_kernel_t<2>, dim3(P.Gr * P.Gc), dim3(MID_THREADS), args, c->mid_smem, c->stream));
    c->launches++;
#ifdef DYB_SERIES_PROF
    {   // diagnostic build: mean / max cycles of each phase over CTAs and terms
        static int calls = 0;
        if (calls++ % 50 == 1) {
            std::vector<long long> h(n_mprof);
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaMemcpy(h.data(), d_mprof, n_mprof * 8, cudaMemcpyDeviceToHost));
            const char* name[8] = {"product", "reduce+publish", "collect", "decision+sums", "update+scalars (thread 0 = scalar warp)", "-", "consume-wait", "loop-gap"};
            const int order[9] = {0, 1, 2, 7, 3, 4, 5, 6, 0};      // stamp order inside a term; the last one is stamp 0 of the next term
            double mean[8] = {0}, mx[8] = {0}, rounds = 0, setup = 0, wait_all = 0, dec = 0, red = 0;
            const int nt = std::min(n, MAX_SERIES_TERMS);
            for (int t = 1; t + 1 < nt; ++t) for (int b = 0; b < mgrid; ++b) {
                const long long* q = &h[((size_t)t * mgrid + b) * 16];
                rounds += double(q[8]); setup += double(q[9] - q[2]); wait_all += double(q[10] - q[7]); dec += double(q[11] - q[10]); red += double(q[12] - q[10]);
                for (int i = 0; i < 8; ++i) {
                    const long long nx = (i < 7) ? q[order[i + 1]] : h[((size_t)(t + 1) * mgrid + b) * 16];
                    const double d = double(nx - q[order[i]]);
                    mean[i] += d; mx[i] = std::max(mx[i], d);
                }
            }
            fprintf(stderr, "mid_prof N=%d grid=%dx%d Cnp=%d NT=%d ST=%d terms=%d:", c->N, P.Gr, P.Gc, P.Cnp, P.NT, P.ST, n);
            for (int i = 0; i < 8; ++i) fprintf(stderr, "  %s mean %.0f max %.0f cyc;", name[i], mean[i] / ((double)std::max(1, nt - 2) * mgrid), mx[i]);
            const double dn = (double)std::max(1, nt - 2) * mgrid;
            fprintf(stderr, "  [thread 0 collect: %.2f polling rounds, %.0f cycles from the publication to the first request; then %.0f waiting for the other threads' words, %.0f for the decision (warp 0), of which %.0f for the combination of the scalars]\n", rounds / dn, setup / dn, wait_all / dn, dec / dn, red / dn);
        }
    }
#endif
    return DYB_OK;
}

// One series in a single cooperative launch when the resident or the mid-size kernel applies.
static bool passes_refgpu(const std::vector<PassParams>& passes) {
    for (const PassParams& pp : passes) if (pp.part[0].test_gpu || pp.part[1].test_gpu) return true;
    return false;
}
static bool single_launch_ok(const dyb_ctx* c, bool refgpu = false) { return resident_ok(c) || mid_ok(c, refgpu); }
static int run_series_single_launch(dyb_ctx* c, const std::vector<PassParams>& passes) {
    return resident_ok(c) ? run_series_resident(c, passes) : run_series_mid(c, passes);
}

static int launch_series_init(dyb_ctx* c, const int adopt[2], const int active[2], int cur, const cplx* sum_scale = nullptr) {
    InitParams I;
    I.M = c->M; I.row0 = c->row0; I.Nc = c->N;
    for (int p = 0; p < 2; ++p) {
        I.adopt[p] = adopt[p]; I.active[p] = active[p];
        I.scale_sum[p] = sum_scale ? 1 : 0;
        I.s_re[p] = sum_scale ? sum_scale[p].real() : 1.0; I.s_im[p] = sum_scale ? sum_scale[p].imag() : 0.0;
    }
    I.psi_b = c->psi_b; I.psi_k = c->psi_k; I.cur_b = c->vb[cur]; I.cur_k = c->vk[cur];
    I.sum_b = c->sum_b; I.sum_k = c->sum_k; I.ctrl = c->ctrl;
    series_init_kernel<<<(2 * c->M + 255) / 256, 256, 0, c->stream>>>(I);
    c->launches++;
    CK(cudaGetLastError());
    if (c->world > 1 && (active[0] || active[1]))     // every rank needs the full starting ket
        CKN(g_nccl.AllGather(c->vk[cur] + (size_t)c->row0 * NQ, c->vk[cur], (size_t)c->M * NQ, ncclDouble, c->comm, c->stream));
    return DYB_OK;
}

static int read_ctrl(dyb_ctx* c) {
    CK(cudaMemcpyAsync(c->h_ctrl, c->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->h_ctrl->peer_timeout)
        return fail(DYB_ECUDA, "row-sharded exchange: a peer's flag did not arrive within the timeout (DYNEMOL_B200_PEER_TIMEOUT_S); "
                               "the ranks are out of step, destroy the contexts");
    return DYB_OK;
}

// ------------------------------------------------------------------------------------------ series coefficients
// Taylor.f:224-239:  c(1) = 1 ; c(k) = -zi * c(k-1) * (tau/(k-1))     (0-based storage here)
static void taylor_coefficient(double tau, cplx* C) {
    const cplx minus_i(0.0, -1.0);
    C[0] = cplx(1.0, 0.0);
    for (int k = 1; k < ORDER; ++k) C[k] = minus_i * C[k - 1] * (tau / (double)k);
}
// Taylor.f:165-171: k_max = first 1-based k >= 2 with |c(k)| < 1e-16, else order
static int taylor_kmax(const cplx* C) {
    for (int k = 2; k <= ORDER; ++k) if (std::abs(C[k - 1]) < 1.0e-16) return k;
    return ORDER;
}

// Taylor_gpu.cpp:553-564 (the reference's GPU variant): k_max = first 0-based k >= 1 with |c_k| < 1e-16, else 25; the series
// then sums the terms k = 1 .. k_max-1, one fewer than Taylor.f whenever the threshold is met.
// Returned 1-based-compatible: the caller uses n_terms = k_ref - 1 for both variants.
static int taylor_kmax_refgpu(const cplx* C) {
    for (int k = 1; k < ORDER; ++k) if (std::abs(C[k]) < 1.0e-16) return k;
    return ORDER;
}

// ------------------------------------------------------------------------------------------ series drivers
// Chebyshev_gpu.cpp:636-643 with the spectral rescaling the reference lacks:
//   R = de*tau ; c_0 = J_0(R) e^{-i ebar tau} ; c_k = 2 (-i)^k J_k(R) e^{-i ebar tau}
static void cheb_coefficient(double tau, double ebar, double de, cplx* C) {
    static const cplx pw[4] = {cplx(1, 0), cplx(0, -1), cplx(-1, 0), cplx(0, 1)};
    const double R = de * tau;
    const cplx ph = std::exp(cplx(0.0, -ebar * tau));
    C[0] = jn(0, R) * ph;
    for (int k = 1; k < ORDER; ++k) C[k] = (2.0 * jn(k, R)) * pw[k & 3] * ph;
}
static double naked_bessel(int n, double x) { return (double)(1 << (n - 2)) * (x * x + 4.0) / pow(x, (double)n); }
// Chebyshev_gpu.cpp:565-574: first k in 6..24 with |c_k nakedBessel(k,R)| < 1e-20, else 25
static int cheb_kmax(const cplx* C, double R) {
    for (int k = 6; k < ORDER; ++k) if (std::abs(C[k] * naked_bessel(k, R)) < 1.0e-20) return k;
    return ORDER;
}

struct Particle {
    bool   present = false, done = true;
    int    phase = 0;                 // 0 first Convergence loop, 1 steady sub-steps, 2 rescale Convergence
    double tau = 0, save_tau = 0, t = 0, norm_ref = 0;
    int    k_ref = 0, n_terms = 0;    // n_terms: dual products of the series being run
    bool   check = false;
    cplx   C[ORDER];
    int    shrinks = 0;
    dyb_trace* tr = nullptr;
};

static void trace_event(dyb_trace* tr, int kind, int k, int ok, double tau) {
    if (!tr) return;
    if (tr->n_events < DYB_MAX_EVENTS) {
        const int e = tr->n_events;
        tr->ev_kind[e] = kind; tr->ev_k[e] = k; tr->ev_ok[e] = ok; tr->ev_tau[e] = tau;
    }
    tr->n_events++;
}

static int compute_norm_ref(dyb_ctx* c, double out[2]) {
    dotc_kernel<<<1, 1024, 0, c->stream>>>(c->M, c->psi_b, c->psi_k + (size_t)c->row0 * NQ, c->scal);
    c->launches++;
    CK(cudaGetLastError());
    if (c->world > 1) CKN(g_nccl.AllReduce(c->scal, c->scal, 4, ncclDouble, ncclSum, c->comm, c->stream));
    CK(cudaMemcpyAsync(c->h_scal, c->scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out[0] = std::abs(cplx(c->h_scal[0], c->h_scal[1]));      // Taylor.f:62
    out[1] = std::abs(cplx(c->h_scal[2], c->h_scal[3]));
    return DYB_OK;
}

// Fill the epilogue parameters of series step s (0-based) for particle q.
//   Taylor    (Taylor.f:182-187 / :90-95):  step s produces term k = s+2:  y = (c_k/c_{k-1}) H' x ; sum += y
//   Chebyshev (Chebyshev_gpu.cpp:552-589):  step s produces phi_j, j = s+1: phi_1 = Ht phi_0, phi_j = 2 Ht phi_{j-1} - phi_{j-2},
//                                           Ht = (H' - ebar)/de ; sum += c_j phi_j ; tests start at j = 2
//   Taylor, reference GPU variant (Taylor_gpu.cpp:566-575): step s produces the raw power Psi_k = H' Psi_{k-1}, k = s+1
//                                           (0-based), sum += c_k Psi_k ; term test of PartPass::test_gpu
static void fill_pass(PartPass& a, const Particle& q, int mode, int s, double ebar, double de, bool refgpu = false) {
    memset(&a, 0, sizeof a);
    a.active = 1; a.norm_ref = q.norm_ref;
    a.last = (s == q.n_terms - 1) ? 1 : 0;
    a.last_ok_by_norm = q.check ? 0 : 1;
    a.test_gpu = refgpu ? 1 : 0;
    if (mode == DYB_MODE_TAYLOR && refgpu) {
        const int k = s + 1;
        a.k = k; a.alpha_re = 1.0; a.scale_term = 1;
        a.c_re = q.C[k].real(); a.c_im = q.C[k].imag();
        a.check_conv = q.check ? 1 : 0;
    } else if (mode == DYB_MODE_TAYLOR) {
        const int k = s + 2;
        const cplx r = q.C[k - 1] / q.C[k - 2];               // Taylor.f:93,185
        a.k = k; a.alpha_re = r.real(); a.alpha_im = r.imag();
        a.check_conv = q.check ? 1 : 0;
    } else {
        const int j = s + 1;
        a.k = j; a.three_term = 1; a.scale_term = 1;
        a.c_re = q.C[j].real(); a.c_im = q.C[j].imag();
        if (j == 1) { a.alpha_re = 1.0 / de; a.beta_re = -ebar / de; a.gamma = 0.0; }
        else        { a.alpha_re = 2.0 / de; a.beta_re = -2.0 * ebar / de; a.gamma = -1.0; }
        a.check_conv = (q.check && j >= 2) ? 1 : 0;
    }
}

// The steady loop of Taylor.f:81-126 when every norm test passes: the tau of each remaining sub-step, computed with the
// very operations the state machine performs (t += tau*h_bar ; a last, shorter sub-step when less than one tau is
// left, :116-121).  Host arithmetic only (exported as dyb_steady_schedule for the CPU tests).
static std::vector<double> steady_schedule(double t, double t_max, double tau, int max_sub) {
    std::vector<double> taus;
    while (t < t_max && (int)taus.size() < max_sub) {
        taus.push_back(tau);
        t += tau * H_BAR;                                         // Taylor.f:116
        if (t_max - t < tau * H_BAR) tau = (t_max - t) / H_BAR;   // Taylor.f:118-121
    }
    return taus;
}

// Propagation(): Taylor.f:35-127 (identical control flow in Chebyshev_gpu.cpp:347-485), one state machine per
// particle, all particles served by the same passes over H'.  Host decisions are taken once per series from
// the device-side control block; per-term decisions (early exit of Convergence) are taken on the device.
static int propagate_series(dyb_ctx* c, int mode_in, double t_init, double t_max, const double* tau_in, double* save_tau, dyb_trace* traces)
{
    // the *_REFGPU modes follow the reference's GPU files decision for decision (Taylor_gpu.cpp:334-480,511-622 /
    // Chebyshev_gpu.cpp:347-485,524-632): same control flow, other term count, term test and place of c_k (fill_pass)
    const bool refgpu = (mode_in == DYB_MODE_TAYLOR_REFGPU || mode_in == DYB_MODE_CHEBYSHEV_REFGPU);
    const int mode = (mode_in == DYB_MODE_TAYLOR || mode_in == DYB_MODE_TAYLOR_REFGPU) ? DYB_MODE_TAYLOR : DYB_MODE_CHEBYSHEV;
    Particle P[2];
    double nref[2];
    int rc = compute_norm_ref(c, nref);
    if (rc) return rc;
    // Chebyshev_gpu.cpp applies the series to H' itself (no spectral rescaling): ebar = 0, de = 1
    const double ebar = (mode_in == DYB_MODE_CHEBYSHEV_REFGPU) ? 0.0 : 0.5 * (c->emax + c->emin);
    const double de   = (mode_in == DYB_MODE_CHEBYSHEV_REFGPU) ? 1.0 : 0.5 * (c->emax - c->emin);
    for (int p = 0; p < c->n_part; ++p) {
        P[p].present = true; P[p].done = false; P[p].phase = 0;
        P[p].tau = tau_in[p]; P[p].norm_ref = nref[p]; P[p].t = t_init;
        P[p].tr = traces ? &traces[p] : nullptr;
        if (P[p].tr) { memset(P[p].tr, 0, sizeof(dyb_trace)); P[p].tr->norm_ref = nref[p]; }
    }
    auto coefficient = [&](Particle& q) {
        if (mode == DYB_MODE_TAYLOR) taylor_coefficient(q.tau, q.C); else cheb_coefficient(q.tau, ebar, de, q.C);
    };
    int adopt[2] = {0, 0};
    long guard = 0;
    c->passes_last = 0;
    for (;;) {
        int active[2] = {0, 0};
        int L = 0;
        cplx sum_scale[2] = {cplx(1, 0), cplx(1, 0)};
        for (int p = 0; p < 2; ++p) {
            Particle& q = P[p];
            if (!q.present || q.done) continue;
            active[p] = 1;
            if (q.phase == 0 || q.phase == 2) {                   // Convergence(): Taylor.f:163-173 / Chebyshev_gpu.cpp:556-575
                coefficient(q);
                q.k_ref = (mode == DYB_MODE_TAYLOR) ? (refgpu ? taylor_kmax_refgpu(q.C) : taylor_kmax(q.C)) : cheb_kmax(q.C, de * q.tau);
                q.check = true;
            } else {                                              // steady sub-step: Taylor.f:90 / Chebyshev_gpu.cpp:418-425
                q.check = false;
            }
            q.n_terms = q.k_ref - 1;                              // Taylor: k = 2..k_ref ; Chebyshev: j = 1..k_ref-1
            if (mode != DYB_MODE_TAYLOR) sum_scale[p] = q.C[0];
            L = std::max(L, q.n_terms);
        }
        if (!active[0] && !active[1]) break;
        if (++guard > 2000000) return fail(DYB_EINVAL, "propagation does not terminate (tau -> 0?)");

        // Small operators (resident kernel) with every active particle in the steady loop: the whole remaining loop of
        // Taylor.f:81-126 goes into ONE launch.  The sub-step schedule (tau, the shortened last sub-step and its
        // coefficients, Taylor.f:116-121) is predicted with the very operations the state machine below performs;
        // the device chains the sub-steps (PartPass::begin / chain) and stops a particle at the first failed norm test.
        bool all_steady = single_launch_ok(c, refgpu) && c->chain_steady;
        for (int p = 0; p < 2; ++p) if (active[p] && P[p].phase != 1) all_steady = false;
        if (all_steady) {
            std::vector<PassParams> passes;
            for (int p = 0; p < 2; ++p) {
                if (!active[p]) continue;
                Particle sim = P[p];
                size_t pos = 0;
                int n_sub = 0;
                const int nt = sim.k_ref - 1;                     // k_ref stays (Taylor.f:118-121 only recomputes the coefficients)
                const int cap = nt >= 1 ? MAX_CHAIN_PASSES / nt : 0;
                for (const double tau_s : steady_schedule(sim.t, t_max, sim.tau, cap)) {
                    if (tau_s != sim.tau) { sim.tau = tau_s; coefficient(sim); }
                    if (passes.size() < pos + nt) {
                        const size_t old_n = passes.size();
                        passes.resize(pos + nt);
                        memset(&passes[old_n], 0, sizeof(PassParams) * (pos + nt - old_n));
                    }
                    sim.check = false; sim.n_terms = nt;
                    for (int j = 0; j < nt; ++j) fill_pass(passes[pos + j].part[p], sim, mode, j, ebar, de, refgpu);
                    if (n_sub > 0) {
                        PartPass& first = passes[pos].part[p];
                        const cplx s0 = (mode == DYB_MODE_TAYLOR) ? cplx(1.0, 0.0) : sim.C[0];
                        first.begin = 1; first.s_re = s0.real(); first.s_im = s0.imag();
                        passes[pos - 1].part[p].chain = 1;
                    }
                    pos += nt; ++n_sub;
                }
                if (n_sub == 0) return fail(DYB_EINVAL, "steady sub-step of %d terms does not fit a launch", sim.k_ref - 1);
            }
            if ((rc = launch_series_init(c, adopt, active, 0, mode == DYB_MODE_TAYLOR ? nullptr : sum_scale))) return rc;
            adopt[0] = adopt[1] = 0;
            if ((rc = run_series_single_launch(c, passes))) return rc;
            if ((rc = read_ctrl(c))) return rc;
            c->passes_last += std::max(active[0] ? c->h_ctrl->part[0].n_terms : 0, active[1] ? c->h_ctrl->part[1].n_terms : 0);
            for (int p = 0; p < 2; ++p) {
                Particle& q = P[p];
                if (!active[p]) continue;
                const PartState& st = c->h_ctrl->part[p];
                const bool failed = st.latched && !st.ok;
                if (q.tr) { q.tr->n_matvec_pairs += st.n_terms; q.tr->last_k_ref = q.k_ref; }
                for (int i = 0; i < st.n_sub_ok; ++i) {           // Taylor.f:102-105, :116-121 for every accepted sub-step
                    if (q.tr) q.tr->n_substeps++;
                    trace_event(q.tr, 2, q.k_ref, 1, q.tau);
                    q.t += q.tau * H_BAR;
                    if (t_max - q.t < q.tau * H_BAR) {
                        q.tau = (t_max - q.t) / H_BAR;
                        coefficient(q);
                    }
                    if (!(q.t < t_max)) q.done = true;
                }
                if (st.n_sub_ok > 0) adopt[p] = 1;                // sum holds the last accepted vector (resident.cuh hands back
                                                                  // the start of a failed sub-step)
                if (failed) {                                     // Taylor.f:108-110
                    if (q.tr) q.tr->n_substeps++;
                    trace_event(q.tr, 2, q.k_ref, 0, q.tau);
                    q.tau *= 0.975;
                    if (q.tr) q.tr->n_rescale++;
                    q.phase = 2;
                }
            }
            continue;
        }

        int prv = 2, cur = 0, nxt = 1;
        if ((rc = launch_series_init(c, adopt, active, cur, mode == DYB_MODE_TAYLOR ? nullptr : sum_scale))) return rc;
        adopt[0] = adopt[1] = 0;
        if (single_launch_ok(c, refgpu) && L <= MAX_SERIES_TERMS) {
            std::vector<PassParams> passes(L);
            for (int s = 0; s < L; ++s) {
                memset(&passes[s], 0, sizeof(PassParams));
                for (int p = 0; p < 2; ++p)
                    if (active[p] && s < P[p].n_terms) fill_pass(passes[s].part[p], P[p], mode, s, ebar, de, refgpu);
            }
            if ((rc = run_series_single_launch(c, passes))) return rc;
        } else
        for (int s = 0; s < L; ++s) {
            EpiParams E = epi_params(c, cur, prv, nxt);
            for (int p = 0; p < 2; ++p) {
                if (active[p] && s < P[p].n_terms) fill_pass(E.pass.part[p], P[p], mode, s, ebar, de, refgpu);
                else E.pass.part[p].active = 0;
            }
            if ((rc = run_term(c, E, cur, nxt, true))) return rc;
            const int old_prv = prv; prv = cur; cur = nxt; nxt = old_prv;
        }
        if ((rc = read_ctrl(c))) return rc;
        c->passes_last += std::max(c->h_ctrl->part[0].latched && active[0] ? c->h_ctrl->part[0].n_terms : 0,
                                   c->h_ctrl->part[1].latched && active[1] ? c->h_ctrl->part[1].n_terms : 0);

        for (int p = 0; p < 2; ++p) {
            Particle& q = P[p];
            if (!active[p]) continue;
            const PartState& st = c->h_ctrl->part[p];
            const bool ok = st.ok != 0;
            if (q.tr) { q.tr->n_matvec_pairs += st.n_terms; q.tr->last_k_ref = q.k_ref; }
            bool advance = false;
            if (q.phase == 0) {                                   // Taylor.f:65-71
                if (q.tr) q.tr->n_convergence_calls++;
                trace_event(q.tr, 1, ok ? st.k_exit : 0, ok, q.tau);
                if (ok) {
                    adopt[p] = 1;
                    q.save_tau = q.tau; save_tau[p] = q.tau;
                    q.t = t_init + q.tau * H_BAR;                 // Taylor.f:73
                    if (t_max - q.t < q.tau * H_BAR) {            // Taylor.f:75-78
                        q.tau = (t_max - q.t) / H_BAR;
                        coefficient(q);
                    }
                    q.phase = 1;
                    if (!(q.t < t_max)) q.done = true;            // Taylor.f:81
                } else {
                    q.tau *= 0.9;
                    if (q.tr) q.tr->n_first_shrink++;
                    if (++q.shrinks > 20000) return fail(DYB_EINVAL, "Convergence never succeeds (tau=%g)", q.tau);
                }
            } else if (q.phase == 1) {                            // Taylor.f:102-114
                if (q.tr) q.tr->n_substeps++;
                trace_event(q.tr, 2, q.k_ref, ok, q.tau);
                if (ok) { adopt[p] = 1; advance = true; }
                else {
                    q.tau *= 0.975;                               // Taylor.f:110
                    if (q.tr) q.tr->n_rescale++;
                    q.phase = 2;
                }
            } else {                                              // rescale loop, Taylor.f:108-113
                if (q.tr) q.tr->n_convergence_calls++;
                trace_event(q.tr, 1, ok ? st.k_exit : 0, ok, q.tau);
                if (ok) { adopt[p] = 1; advance = true; q.phase = 1; }
                else {
                    q.tau *= 0.975;
                    if (q.tr) q.tr->n_rescale++;
                    if (++q.shrinks > 20000) return fail(DYB_EINVAL, "rescaling tau never converges (tau=%g)", q.tau);
                }
            }
            if (advance) {
                q.t += q.tau * H_BAR;                             // Taylor.f:116
                if (t_max - q.t < q.tau * H_BAR) {                // Taylor.f:118-121
                    q.tau = (t_max - q.t) / H_BAR;
                    coefficient(q);
                }
                if (!(q.t < t_max)) q.done = true;
            }
        }
    }
    // adopt the last accepted sums
    if (adopt[0] || adopt[1]) {
        const int none[2] = {0, 0};
        if ((rc = launch_series_init(c, adopt, none, 0))) return rc;
    }
    CK(cudaStreamSynchronize(c->stream));
    for (int p = 0; p < c->n_part; ++p) if (P[p].tr) P[p].tr->final_tau = P[p].tau;
    return DYB_OK;
}

// ------------------------------------------------------------------------------------------ single-expansion Chebyshev
// DYB_MODE_CHEBYSHEV_FULL: ONE Chebyshev expansion for the whole interval t_init .. t_max of a nuclear step instead of the
// reference's chain of order-25 series (Chebyshev_gpu.cpp:347-485 caps the order at 25, so a 0.5 fs step at R = dE*tau
// ~ 500 is cut into ~95 sub-steps of 24 terms; the Bessel coefficients only start to decay at k ~ R, which makes a
// single expansion of R + O(R^(1/3)) terms the cheaper AND the more accurate way to cross the interval):
//     psi(t_max) = sum_{k<K} c_k T_k(Ht) psi ,  c_0 = J_0(R) e^{-i ebar tau} , c_k = 2 (-i)^k J_k(R) e^{-i ebar tau}   (same series,
//     Chebyshev_gpu.cpp:552-589,636-643), K = first k > R with 2|J_k(R)| < 1e-15 (beyond k = R the coefficients fall
//     super-exponentially, so the tail is below that bound too).
// One norm test at the end (the reference's 1e-8, Taylor.f:104).  A failed test means the spectral interval did not
// enclose the spectrum (T_k grows outside [-1,1]) or the order is out of reach: the interval is widened once, then the
// step is cut in two, keeping the packets of the last accepted expansion.
static std::vector<cplx> cheb_full_coefficients(double tau, double ebar, double de) {
    static const cplx pw[4] = {cplx(1, 0), cplx(0, -1), cplx(-1, 0), cplx(0, 1)};
    const double R = de * tau;
    const cplx ph = std::exp(cplx(0.0, -ebar * tau));
    std::vector<cplx> C;
    C.push_back(jn(0, R) * ph);
    for (int k = 1; k < (1 << 20); ++k) {
        const double j = jn(k, R);
        if ((double)k > R && k >= 2 && 2.0 * fabs(j) < 1.0e-15) break;
        C.push_back((2.0 * j) * pw[k & 3] * ph);
    }
    return C;
}

// the terms of one expansion through whichever path applies: the resident kernel in ONE launch, else a dual product +
// fused epilogue per term (PDL-chained), with the three rotating vectors of the recurrence
static int run_pass_list(dyb_ctx* c, const std::vector<PassParams>& passes) {
    int rc;
    if (single_launch_ok(c, passes_refgpu(passes)) && (int)passes.size() <= MAX_CHAIN_PASSES) return run_series_single_launch(c, passes);
    int prv = 2, cur = 0, nxt = 1;
    for (const PassParams& pp : passes) {
        EpiParams E = epi_params(c, cur, prv, nxt);
        E.pass = pp;
        if ((rc = run_term(c, E, cur, nxt, true))) return rc;
        const int old_prv = prv; prv = cur; cur = nxt; nxt = old_prv;
    }
    return DYB_OK;
}

static int propagate_cheb_full(dyb_ctx* c, double t_init, double t_max, const double* tau_in, double* save_tau, dyb_trace* traces)
{
    double nref[2];
    int rc = compute_norm_ref(c, nref);
    if (rc) return rc;
    for (int p = 0; p < c->n_part; ++p) if (traces) { memset(&traces[p], 0, sizeof(dyb_trace)); traces[p].norm_ref = nref[p]; }
    // Taylor.f:65-81: a slice of zero (or negative) length still advances by the given tau
    double remaining = (t_max - t_init) / H_BAR;
    if (!(remaining > 0.0)) remaining = std::max(tau_in[0], c->n_part > 1 ? tau_in[1] : tau_in[0]);
    double emin = c->emin, emax = c->emax;
    double tau = remaining;
    int adopt[2] = {0, 0};
    const int active[2] = {1, c->n_part > 1 ? 1 : 0};
    bool first_ok = false, widened = false;
    c->passes_last = 0;
    for (int attempt = 0; remaining > 0.0; ++attempt) {
        if (attempt > 64) return fail(DYB_EINVAL, "single-expansion Chebyshev step does not pass the norm test (tau=%g)", tau);
        tau = std::min(tau, remaining);
        const double ebar = 0.5 * (emax + emin), de = 0.5 * (emax - emin);
        std::vector<cplx> C = cheb_full_coefficients(tau, ebar, de);
        if (single_launch_ok(c) && (int)C.size() - 1 > MAX_CHAIN_PASSES) { tau *= 0.5; continue; }    // one launch holds 4096 terms
        const int K = (int)C.size();
        if (K < 2) C.push_back(cplx(0.0, 0.0));
        const int n_terms = std::max(1, K - 1);
        std::vector<PassParams> passes(n_terms);
        for (int s = 0; s < n_terms; ++s) {
            memset(&passes[s], 0, sizeof(PassParams));
            for (int p = 0; p < 2; ++p) {
                if (!active[p]) continue;
                PartPass& a = passes[s].part[p];
                const int j = s + 1;
                a.active = 1; a.norm_ref = nref[p]; a.k = j; a.three_term = 1; a.scale_term = 1;
                a.c_re = C[j].real(); a.c_im = C[j].imag();
                if (j == 1) { a.alpha_re = 1.0 / de; a.beta_re = -ebar / de; a.gamma = 0.0; }
                else        { a.alpha_re = 2.0 / de; a.beta_re = -2.0 * ebar / de; a.gamma = -1.0; }
                a.last = (s == n_terms - 1) ? 1 : 0; a.last_ok_by_norm = 1;
            }
        }
        const cplx sum_scale[2] = {C[0], C[0]};
        if ((rc = launch_series_init(c, adopt, active, 0, sum_scale))) return rc;
        adopt[0] = adopt[1] = 0;
        if ((rc = run_pass_list(c, passes))) return rc;
        if ((rc = read_ctrl(c))) return rc;
        c->passes_last += n_terms;
        bool ok = true;
        for (int p = 0; p < c->n_part; ++p) {
            const PartState& st = c->h_ctrl->part[p];
            ok = ok && st.latched && st.ok;
            if (traces) {
                traces[p].n_convergence_calls++; traces[p].n_matvec_pairs += n_terms; traces[p].last_k_ref = K;
                trace_event(&traces[p], 1, st.ok ? K : 0, st.ok, tau);
            }
        }
        if (ok) {
            adopt[0] = active[0]; adopt[1] = active[1];
            if (!first_ok) { first_ok = true; for (int p = 0; p < c->n_part; ++p) save_tau[p] = tau; }
            remaining -= tau;
            if (remaining < 1e-12 * tau) remaining = 0.0;
        } else if (!widened) {                                  // most likely cause: an eigenvalue outside the estimated interval
            const double w = emax - emin;
            emin -= 0.10 * w; emax += 0.10 * w; widened = true;
            if (traces) for (int p = 0; p < c->n_part; ++p) traces[p].n_rescale++;
        } else {
            tau *= 0.5;
            if (traces) for (int p = 0; p < c->n_part; ++p) traces[p].n_rescale++;
        }
    }
    if (adopt[0] || adopt[1]) {
        const int none[2] = {0, 0};
        if ((rc = launch_series_init(c, adopt, none, 0))) return rc;
    }
    CK(cudaStreamSynchronize(c->stream));
    if (traces) for (int p = 0; p < c->n_part; ++p) traces[p].final_tau = tau;
    return DYB_OK;
}

// ------------------------------------------------------------------------------------------ C ABI: native
extern "C" {

const char* dyb_last_error(void) { return g_err.c_str(); }
const char* dyb_version(void) { return "dynemol_b200 0.1 (sm_100a)"; }

// Host-only: the launch plan for an N-column, n_rows-row operator on a GPU with sm_count SMs (no device needed).
// out8 = {panels, tiles_per_panel, tiles, grid, segments, tile_cols, panel_rows, Ncpad}; seg_base[grid] and
// pseg_start[panels+1] are filled when non-NULL (sized by the caller from a first call).
int dyb_plan(int N, int n_rows, int sm_count, int64_t* out8, int32_t* seg_base, int32_t* pseg_start) {
    if (N <= 0 || n_rows <= 0 || sm_count <= 0 || !out8) return fail(DYB_EINVAL, "bad argument");
    const Plan p = make_plan(N, n_rows, sm_count);
    out8[0] = p.NP; out8[1] = p.TPP; out8[2] = p.T; out8[3] = p.grid; out8[4] = p.n_seg; out8[5] = TILE_COLS; out8[6] = PANEL_ROWS; out8[7] = p.Ncpad;
    if (seg_base) for (int b = 0; b < p.grid; ++b) seg_base[b] = p.seg_base[b];
    if (pseg_start) for (int q = 0; q <= p.NP; ++q) pseg_start[q] = p.pseg_start[q];
    return DYB_OK;
}

// Host-only: the blocking of the shared-memory-resident series kernel (resident.cuh) for an N x N operator on a GPU
// with sm_count SMs and smem_optin bytes of opt-in shared memory per block.  out6 = {grid side Gd, block size Bs,
// smem column stride, dynamic smem bytes, threads per CTA, fits (0/1)}.
struct ResidentPlan { int Gd, Bs, ldS; size_t smem; bool fits; };
static ResidentPlan make_resident_plan(int N, int sm_count, size_t smem_optin, size_t static_smem) {
    ResidentPlan r;
    int gd_max = 1;
    while ((gd_max + 1) * (gd_max + 1) <= sm_count && gd_max + 1 <= RES_MAX_GD) ++gd_max;          // 12 on 148 SMs
    r.Gd = std::min(gd_max, std::max(1, N / 32));                                                  // blocks of >= 32 rows
    r.Bs = (N + r.Gd - 1) / r.Gd;
    r.ldS = r.Bs | 1;
    const ResidentSmem L(r.Bs, r.ldS);
    r.smem = L.bytes();
    r.fits = r.Bs <= RES_MAX_BS && r.smem <= (size_t)RES_SMEM_MAX && r.smem + static_smem <= smem_optin;
    return r;
}
int dyb_resident_plan(int N, int sm_count, int64_t smem_optin, int64_t* out6) {
    if (N <= 0 || sm_count <= 0 || smem_optin <= 0 || !out6) return fail(DYB_EINVAL, "bad argument");
    const ResidentPlan r = make_resident_plan(N, sm_count, (size_t)smem_optin, 2048);
    out6[0] = r.Gd; out6[1] = r.Bs; out6[2] = r.ldS; out6[3] = (int64_t)r.smem; out6[4] = RES_THREADS; out6[5] = r.fits ? 1 : 0;
    return DYB_OK;
}

// Host-only: the blocking of the streamed one-launch series kernel (mid.cuh).  out12 = {block rows R, tile columns TC, grid rows
// Gr, grid columns Gc, block columns Cnp, tiles per term, ring stages, indices per owner E, owner CTAs, words an owner collects
// per term (negative: the collect table is the 16-bit one), dynamic smem bytes, fits (0/1)}.
int dyb_mid_plan(int N, int sm_count, int64_t smem_optin, int64_t* out12) {
    if (N <= 0 || sm_count <= 0 || smem_optin <= 0 || !out12) return fail(DYB_EINVAL, "bad argument");
    const MidPlan m = make_mid_plan(N, sm_count, (size_t)smem_optin, 2048);
    out12[0] = m.R; out12[1] = m.TC; out12[2] = m.Gr; out12[3] = m.Gc; out12[4] = m.Cnp; out12[5] = m.NT; out12[6] = m.ST;
    out12[7] = m.E; out12[8] = m.n_own; out12[9] = (int64_t)(m.Gr + m.Gc) * m.E * NQ * (m.tab16 ? -1 : 1);
    out12[10] = (int64_t)m.smem; out12[11] = m.fits ? 1 : 0;
    return DYB_OK;
}

// Host-only: the 25 series coefficients the library uses for a given tau and the number of terms k_max it would sum
// (Taylor.f:224-239 + :165-171 ; Chebyshev_gpu.cpp:636-643 + :565-574 on the interval ebar +- de).  For the CPU tests.
int dyb_series_coefficients(int mode, double tau, double ebar, double de, dyb_complex* out25, int* k_max) {
    if (!out25 || (mode != DYB_MODE_TAYLOR && mode != DYB_MODE_CHEBYSHEV)) return fail(DYB_EINVAL, "bad argument");
    cplx C[ORDER];
    int km;
    if (mode == DYB_MODE_TAYLOR) { taylor_coefficient(tau, C); km = taylor_kmax(C); }
    else {
        if (!(de > 0.0)) return fail(DYB_EINVAL, "Chebyshev mode needs a spectral half width de > 0");
        cheb_coefficient(tau, ebar, de, C); km = cheb_kmax(C, de * tau);
    }
    for (int k = 0; k < ORDER; ++k) { out25[k].re = C[k].real(); out25[k].im = C[k].imag(); }
    if (k_max) *k_max = km;
    return DYB_OK;
}

// Host-only: tau of every remaining steady sub-step of one particle (see steady_schedule above).  Returns the count.
int dyb_steady_schedule(double t, double t_max, double tau, int max_sub, double* out_tau) {
    if (max_sub < 0 || (max_sub > 0 && !out_tau)) return -1;
    const std::vector<double> v = steady_schedule(t, t_max, tau, max_sub);
    for (size_t i = 0; i < v.size(); ++i) out_tau[i] = v[i];
    return (int)v.size();
}

int dyb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int dyb_destroy(dyb_ctx* c) {
    if (!c) return DYB_OK;
    cudaSetDevice(c->device);
    if (c->out_thread.joinable()) c->out_thread.join();
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_h_up) cudaEventDestroy(c->ev_h_up);
    if (c->ev_H_done) cudaEventDestroy(c->ev_H_done);
    if (c->vk_in_comm) c->vk[0] = c->vk[1] = c->vk[2] = nullptr;       // they live inside comm_buf, freed below
    double** bufs[] = {&c->H, &c->S, &c->psi_b, &c->psi_k, &c->sum_b, &c->sum_k, &c->vb[0], &c->vb[1], &c->vb[2],
                       &c->vk[0], &c->vk[1], &c->vk[2], &c->ket_slab, &c->bra_slab, &c->blockpart, &c->scal, &c->io};
    for (auto b : bufs) if (*b) cudaFree(*b);
    if (c->p2p && !c->p2p_local) for (int r = 0; r < c->world; ++r) if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->comm_buf) cudaFree(c->comm_buf);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (double** b : {&c->rs_send, &c->rs_recv, &c->scal_all, &c->full_tmp}) if (*b) cudaFree(*b);
    if (c->ipiv) cudaFree(c->ipiv);
    if (c->colblk) cudaFree(c->colblk);
    for (double* b : {c->ehr_A, c->ehr_X, c->ehr_K, c->ehr_vec}) if (b) cudaFree(b);
    for (double* b : {c->lz_V, c->lz_W, c->lz_dots}) if (b) cudaFree(b);
    if (c->lz_state) cudaFree(c->lz_state);
    if (c->seg_base) cudaFree(c->seg_base);
    if (c->pseg_start) cudaFree(c->pseg_start);
    if (c->frag) cudaFree(c->frag);
    if (c->ctrl) cudaFree(c->ctrl);
    if (c->d_passes) cudaFree(c->d_passes);
    if (c->gbar) cudaFree(c->gbar);
    if (c->res_pk) cudaFree(c->res_pk);
    if (c->res_pb) cudaFree(c->res_pb);
    if (c->res_dscal) cudaFree(c->res_dscal);
    if (c->res_psi) cudaFree(c->res_psi);
    for (double* b : {c->mid_pk, c->mid_pb, c->mid_xx, c->mid_sc}) if (b) cudaFree(b);
    if (c->h_ctrl) cudaFreeHost(c->h_ctrl);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    for (auto e : c->ev) cudaEventDestroy(e);
    if (c->sparams) cusolverDnDestroyParams(c->sparams);
    if (c->solver) cusolverDnDestroy(c->solver);
    if (c->blas) cublasDestroy(c->blas);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return DYB_OK;
}

int dyb_create(dyb_ctx** out, int device, int N, int row0, int n_rows) {
    if (!out) return fail(DYB_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = ensure_device(device);
    if (rc) return rc;
    if (N <= 0 || row0 < 0 || n_rows <= 0 || row0 + n_rows > N) return fail(DYB_EINVAL, "bad shape N=%d row0=%d n_rows=%d", N, row0, n_rows);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(DYB_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    dyb_ctx* c = new dyb_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount;
    c->N = N; c->row0 = row0; c->M = n_rows;
    c->ld = ((long long)n_rows + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
#define CKC(x) do { int r_ = (x); if (r_) { dyb_destroy(c); return r_; } } while (0)
#define CKCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int r_ = fail(e_ == cudaErrorMemoryAllocation ? DYB_ENOMEM : DYB_ECUDA, \
        "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); dyb_destroy(c); return r_; } } while (0)
    CKCU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CKCU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CKCU(cudaEventCreateWithFlags(&c->ev_h_up, cudaEventDisableTiming));
    CKCU(cudaEventCreateWithFlags(&c->ev_H_done, cudaEventDisableTiming));
    CKC(build_plan(c));
    c->Lq = (size_t)std::max(c->NP * PANEL_ROWS, c->Ncpad) + PANEL_ROWS;
    CKCU(cudaMalloc(&c->H, (size_t)c->ld * N * sizeof(double)));
    CKCU(cudaMemsetAsync(c->H, 0, (size_t)c->ld * N * sizeof(double), c->stream));
    CKC(alloc_zero(&c->psi_b, c->Lq * NQ)); CKC(alloc_zero(&c->psi_k, c->Lq * NQ));
    CKC(alloc_zero(&c->sum_b, c->Lq * NQ)); CKC(alloc_zero(&c->sum_k, c->Lq * NQ));
    for (int i = 0; i < 3; ++i) { CKC(alloc_zero(&c->vb[i], c->Lq * NQ)); CKC(alloc_zero(&c->vk[i], c->Lq * NQ)); }
    CKC(alloc_zero(&c->ket_slab, (size_t)std::max(1, c->n_seg) * PANEL_ROWS * NQ));
    CKC(alloc_zero(&c->bra_slab, (size_t)c->NP * c->Ncpad * NQ));
    CKC(alloc_zero(&c->blockpart, (size_t)epi_grid(c) * EPI_NS_RG));
    CKC(alloc_zero(&c->scal, 64));
    CKC(alloc_zero(&c->io, (size_t)N * 4 * 2));            // staging for host <-> quad conversion / S^-1 solves
    CKCU(cudaMalloc(&c->ctrl, sizeof(Ctrl)));
    CKCU(cudaMemset(c->ctrl, 0, sizeof(Ctrl)));
    CKCU(cudaMallocHost(&c->h_ctrl, sizeof(Ctrl)));
    CKCU(cudaMallocHost(&c->h_scal, 128 * sizeof(double)));
    CKC(build_tensor_map(c));
    CKCU(cudaFuncSetAttribute(dual_matvec_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TmaSmem::total));
    CKCU(cudaMalloc(&c->d_passes, sizeof(PassParams) * MAX_CHAIN_PASSES));
    CKCU(cudaMalloc(&c->gbar, sizeof(unsigned long long)));
    if (row0 == 0 && n_rows == N) {    // resident.cuh: does a Gd x Gd blocking of H' fit the shared memories?
        int smem_optin = 0;
        CKCU(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        cudaFuncAttributes fa;
        CKCU(cudaFuncGetAttributes(&fa, resident_series_kernel_t<true>));
        const ResidentPlan rp = make_resident_plan(N, c->sm_count, (size_t)smem_optin, fa.sharedSizeBytes);
        const int Gd = rp.Gd, Bs = rp.Bs;
        bool launchable = rp.fits;
        if (launchable) {   // a cooperative grid must be co-resident: one CTA per SM, and the device must support it
            CKCU(cudaFuncSetAttribute(resident_series_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_SMEM_MAX));
            CKCU(cudaFuncSetAttribute(resident_series_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_SMEM_MAX));
            int coop = 0, per_sm = 0;
            CKCU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
            CKCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resident_series_kernel_t<true>, RES_THREADS, rp.smem));
            launchable = coop && (long)per_sm * c->sm_count >= (long)Gd * Gd;
        }
        if (launchable) {
            c->res_Gd = Gd; c->res_Bs = Bs; c->res_ldS = rp.ldS; c->res_smem = rp.smem;
            CKC(alloc_zero(&c->res_pk, (size_t)2 * Gd * Gd * Bs * NQ)); CKC(alloc_zero(&c->res_pb, (size_t)2 * Gd * Gd * Bs * NQ));
            CKC(alloc_zero(&c->res_dscal, (size_t)2 * Gd * RES_NS));
            CKC(alloc_zero(&c->res_psi, (size_t)Gd * Gd * RES_THREADS * 2));
        }
    }
    if (row0 == 0 && n_rows == N && c->ld % MID_SUB == 0) {    // mid.cuh: blocking, buffers, tensor map of the streamed one-launch kernel
        int smem_optin = 0, coop = 0, per_sm = 0;
        CKCU(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        CKCU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        cudaFuncAttributes fa;
        CKCU(cudaFuncGetAttributes(&fa, mid_series_kernel_t<2>));
        const MidPlan mp = make_mid_plan(N, c->sm_count, (size_t)smem_optin, fa.sharedSizeBytes);
        bool launchable = mp.fits && coop;
        if (launchable) {
            CKCU(cudaFuncSetAttribute(mid_series_kernel_t<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_SMEM_MAX));
            CKCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mid_series_kernel_t<2>, MID_THREADS, mp.smem));
            launchable = (long)per_sm * c->sm_count >= (long)mp.Gr * mp.Gc;
        }
        if (launchable) {
            if (const char* e = getenv("DYNEMOL_B200_MID_MAX")) c->mid_auto_max = atoi(e);
            if (const char* e = getenv("DYNEMOL_B200_MID_L2MB")) c->mid_l2_mb = atof(e);
            MidParams& P = c->mid_P;
            memset(&P, 0, sizeof P);
            P.N = N; P.Gr = mp.Gr; P.Gc = mp.Gc; P.Cnp = mp.Cnp; P.NT = mp.NT; P.ST = mp.ST;
            P.E = mp.E; P.n_own = mp.n_own; P.tab16 = mp.tab16;
            const double bytes = 8.0 * (double)c->ld * N;
            P.l2_frac = c->mid_l2_mb <= 0.0 ? 0.f : (float)std::min(1.0, c->mid_l2_mb * 1.0e6 / bytes);
            c->mid_smem = mp.smem;
            CKC(build_tensor_map_mid(c, mp.WR, mp.TC));
            const size_t G = (size_t)mp.Gr * mp.Gc;
            // two doubles of storage per published double (epoch-tagged 16-byte words); zero = epoch 0 = never valid
            const size_t nd[4] = {2 * 2 * G * mp.R * NQ, 2 * 2 * G * mp.Cnp * NQ, 2 * (size_t)2 * 2 * N * NQ, 2 * 4 * G * 8};
            CKC(alloc_zero(&c->mid_pk, nd[0])); CKC(alloc_zero(&c->mid_pb, nd[1]));
            CKC(alloc_zero(&c->mid_xx, nd[2])); CKC(alloc_zero(&c->mid_sc, nd[3]));
            for (int i = 0; i < 4; ++i) c->mid_bytes[i] = nd[i] * 8;
            c->mid_fits = true;
        }
    }
    if (const char* e = getenv("DYNEMOL_B200_CHAIN")) c->chain_steady = (e[0] != '0');
    if (const char* e = getenv("DYNEMOL_B200_SERIES")) {
        c->series_kind = !strcmp(e, "term") ? DYB_SERIES_PER_TERM : !strcmp(e, "resident") ? DYB_SERIES_RESIDENT
                       : !strcmp(e, "mid") ? DYB_SERIES_MID : DYB_SERIES_AUTO;
    }
    CKCU(cudaDeviceSynchronize());     // the zero fills above ran on the legacy stream; c->stream is non-blocking
#undef CKC
#undef CKCU
    if (const char* e = getenv("DYNEMOL_B200_PDL")) c->use_pdl = (e[0] != '0');
    *out = c;
    return DYB_OK;
}

int dyb_set_kernel(dyb_ctx* c, int v) {
    if (!c) return fail(DYB_EINVAL, "ctx is NULL");
    if (v == DYB_KERNEL_AUTO) v = DYB_KERNEL_TMA;
    if (v != DYB_KERNEL_TMA && v != DYB_KERNEL_LDG) return fail(DYB_EINVAL, "unknown kernel variant %d", v);
    c->variant = v;
    return DYB_OK;
}

int dyb_set_series_kernel(dyb_ctx* c, int kind) {
    if (!c) return fail(DYB_EINVAL, "ctx is NULL");
    if (kind != DYB_SERIES_AUTO && kind != DYB_SERIES_PER_TERM && kind != DYB_SERIES_RESIDENT && kind != DYB_SERIES_MID)
        return fail(DYB_EINVAL, "unknown series kernel %d", kind);
    if (kind == DYB_SERIES_MID && !c->mid_fits) return fail(DYB_EINVAL, "the mid-size series kernel does not apply to this context (N=%d, rows=%d)", c->N, c->M);
    c->series_kind = kind;
    return DYB_OK;
}

int dyb_get_info(dyb_ctx* c, int64_t* o) {
    if (!c || !o) return fail(DYB_EINVAL, "NULL argument");
    memset(o, 0, 16 * sizeof(int64_t));
    o[0] = c->N; o[1] = c->ld; o[2] = c->M; o[3] = c->grid; o[4] = c->T; o[5] = c->n_seg; o[6] = c->sm_count;
    o[7] = TmaSmem::total; o[8] = c->variant; o[9] = c->NP; o[10] = c->TPP; o[11] = c->passes_last; o[12] = c->p2p ? 1 : 0;
    o[13] = resident_ok(c) ? DYB_SERIES_RESIDENT : mid_ok(c, false) ? DYB_SERIES_MID : DYB_SERIES_PER_TERM;
    o[14] = c->res_Gd; o[15] = c->res_Bs;
    return DYB_OK;
}

int dyb_sync(dyb_ctx* c) {
    if (!c) return fail(DYB_EINVAL, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

int64_t dyb_launch_count(dyb_ctx* c) { return c ? c->launches : 0; }

// ---- operator ------------------------------------------------------------------------------------
int dyb_upload_hprime(dyb_ctx* c, const double* h_H, int64_t lda) {
    if (!c || !h_H || lda < c->N) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy2DAsync(c->H, (size_t)c->ld * 8, h_H + c->row0, (size_t)lda * 8, (size_t)c->M * 8, c->N,
                         cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->have_factor = false;
    return DYB_OK;
}

int dyb_upload_hprime_device(dyb_ctx* c, const void* d_H, int64_t lda) {
    if (!c || !d_H || lda < c->N) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy2DAsync(c->H, (size_t)c->ld * 8, reinterpret_cast<const double*>(d_H) + c->row0, (size_t)lda * 8,
                         (size_t)c->M * 8, c->N, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->have_factor = false;
    return DYB_OK;
}

int dyb_upload_hprime_rows_device(dyb_ctx* c, const void* d_rows, int64_t lda, int local_row0, int n_rows) {
    if (!c || !d_rows || n_rows < 1 || lda < n_rows || local_row0 < 0 || local_row0 + n_rows > c->M) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy2DAsync(c->H + local_row0, (size_t)c->ld * 8, d_rows, (size_t)lda * 8, (size_t)n_rows * 8, c->N, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->have_factor = false;
    return DYB_OK;
}

int dyb_hprime_device(dyb_ctx* c, void** d_ptr, int64_t* ld) {
    if (!c || !d_ptr || !ld) return fail(DYB_EINVAL, "NULL argument");
    *d_ptr = c->H; *ld = c->ld;
    return DYB_OK;
}

int dyb_download_hprime(dyb_ctx* c, double* h_H, int64_t lda) {
    if (!c || !h_H || lda < c->M) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy2DAsync(h_H, (size_t)lda * 8, c->H, (size_t)c->ld * 8, (size_t)c->M * 8, c->N, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

int dyb_download_hprime_rows_device(dyb_ctx* c, void* d_dst, int64_t ldd, int local_row0, int n_rows) {
    if (!c || !d_dst || n_rows < 1 || ldd < n_rows || local_row0 < 0 || local_row0 + n_rows > c->M) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpy2DAsync(d_dst, (size_t)ldd * 8, c->H + local_row0, (size_t)c->ld * 8, (size_t)n_rows * 8, c->N, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

static int ensure_solver(dyb_ctx* c) {
    if (!c->solver) {
        CKS(cusolverDnCreate(&c->solver));
        CKS(cusolverDnSetStream(c->solver, c->stream));
        CKS(cusolverDnCreateParams(&c->sparams));
    }
    if (!c->S) CK(cudaMalloc(&c->S, (size_t)c->N * c->N * sizeof(double)));
    return DYB_OK;
}

// H' = S^-1 h with S in c->S (N x N, lda N, destroyed) and h already in c->H (ld).
// ElHl_Chebyshev.f:206-210 computes inv(S) (dsytrf/dsytri) and then dsymm; here S is factorised once
// (Cholesky; LU with partial pivoting if S is not numerically SPD, as the reference's GPU flavour
// GPU_Interface.cpp:910-929) and the N right-hand sides are solved in place: fewer flops, no explicit
// inverse, same result to O(cond(S) eps).  The factor is kept for AO_bra = S^-1 Psi_bra.
struct DevScratch {                         // freed on every exit path
    void* p = nullptr;
    ~DevScratch() { if (p) cudaFree(p); }
};

// `load_h` brings h into c->H; it runs after the factorisation has been queued, so an upload on the copy stream (or a
// host-blocking pageable copy) overlaps potrf.  `reload_S` restores S for the LU fallback.
static int factor_and_solve(dyb_ctx* c, const std::function<int()>& reload_S, const std::function<int()>& load_h = nullptr) {
    const int64_t n = c->N;
    c->have_factor = false;
    size_t wd = 0, wh = 0;
    int* d_info = reinterpret_cast<int*>(c->scal + 32);
    CKS(cusolverDnXpotrf_bufferSize(c->solver, c->sparams, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, c->S, n, CUDA_R_64F, &wd, &wh));
    DevScratch w1; std::vector<char> h_work(wh ? wh : 1);
    if (wd) CK(cudaMalloc(&w1.p, wd));
    cusolverStatus_t st = cusolverDnXpotrf(c->solver, c->sparams, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, c->S, n, CUDA_R_64F,
                                           w1.p, wd, h_work.data(), wh, d_info);
    int info = 0;
    if (st == CUSOLVER_STATUS_SUCCESS) CK(cudaMemcpyAsync(c->h_scal + 100, d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (load_h) { int rc = load_h(); if (rc) return rc; }
    if (st == CUSOLVER_STATUS_SUCCESS) { CK(cudaStreamSynchronize(c->stream)); info = *reinterpret_cast<int*>(c->h_scal + 100); }
    if (st != CUSOLVER_STATUS_SUCCESS) return fail(DYB_ECUDA, "cusolverDnXpotrf status %d", (int)st);
    if (info == 0) {
        CKS(cusolverDnXpotrs(c->solver, c->sparams, CUBLAS_FILL_MODE_UPPER, n, n, CUDA_R_64F, c->S, n, CUDA_R_64F, c->H, c->ld, d_info));
        c->factor_is_lu = false;
    } else {
        // S is not numerically positive definite: LU with partial pivoting, the route of the reference's GPU
        // flavour (magma_dgetrf_gpu/dgetri_gpu, GPU_Interface.cpp:910-929).  potrf destroyed part of S: reload it.
        int rc = reload_S();
        if (rc) return rc;
        if (!c->ipiv) CK(cudaMalloc(&c->ipiv, sizeof(int64_t) * n));
        CKS(cusolverDnXgetrf_bufferSize(c->solver, c->sparams, n, n, CUDA_R_64F, c->S, n, CUDA_R_64F, &wd, &wh));
        DevScratch w2; h_work.resize(wh ? wh : 1);
        if (wd) CK(cudaMalloc(&w2.p, wd));
        st = cusolverDnXgetrf(c->solver, c->sparams, n, n, CUDA_R_64F, c->S, n, c->ipiv, CUDA_R_64F, w2.p, wd, h_work.data(), wh, d_info);
        if (st == CUSOLVER_STATUS_SUCCESS) { CK(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
        if (st != CUSOLVER_STATUS_SUCCESS) return fail(DYB_ECUDA, "cusolverDnXgetrf status %d", (int)st);
        if (info != 0) return fail(DYB_ESINGULAR, "overlap matrix is singular (getrf info=%d)", info);
        CKS(cusolverDnXgetrs(c->solver, c->sparams, CUBLAS_OP_N, n, n, CUDA_R_64F, c->S, n, c->ipiv, CUDA_R_64F, c->H, c->ld, d_info));
        c->factor_is_lu = true;
    }
    c->have_factor = true;
    return DYB_OK;
}

// Wait for the download of H' started by dyb_form_hprime_async (no-op when none is pending).
int dyb_wait_outputs(dyb_ctx* c) {
    if (!c) return fail(DYB_EINVAL, "ctx is NULL");
    if (c->out_thread.joinable()) {
        c->out_thread.join();
        if (c->out_rc) return fail(c->out_rc, "%s", c->out_err.c_str());
    }
    return DYB_OK;
}

// H' = S^-1 h from host S, h with the transfers overlapped: S goes up first and its factorisation is queued; h goes up on
// the copy stream while potrf runs; the download of H' (8 N^2 bytes) starts on the copy stream as soon as the solve ends
// and runs beside whatever the caller queues next (Lanczos, the series).  It is issued from a helper thread because a
// copy into pageable host memory blocks the issuing thread; with pinned buffers (GPU_Pin, ElHl_Chebyshev_GPU.f:109-111)
// it is a plain asynchronous DMA.  dyb_wait_outputs() joins it: h_H_out is complete only after that call.
int dyb_form_hprime_async(dyb_ctx* c, const double* h_S, const double* h_h, double* h_H_out) {
    if (!c || !h_S || !h_h) return fail(DYB_EINVAL, "NULL argument");
    if (c->M != c->N) return fail(DYB_EINVAL, "dyb_form_hprime needs the full matrix on one device (row shard given)");
    CK(cudaSetDevice(c->device));
    int rc = dyb_wait_outputs(c);
    if (rc) return rc;
    if ((rc = ensure_solver(c))) return rc;
    const size_t n = c->N;
    auto load_S = [&]() -> int { CK(cudaMemcpyAsync(c->S, h_S, n * n * 8, cudaMemcpyHostToDevice, c->stream)); return DYB_OK; };
    auto load_h = [&]() -> int {
        CK(cudaMemcpy2DAsync(c->H, (size_t)c->ld * 8, h_h, n * 8, n * 8, n, cudaMemcpyHostToDevice, c->copy_stream));
        CK(cudaEventRecord(c->ev_h_up, c->copy_stream));
        CK(cudaStreamWaitEvent(c->stream, c->ev_h_up, 0));
        return DYB_OK;
    };
    CK(cudaStreamSynchronize(c->copy_stream));            // nothing of a previous call may still be reading c->H
    if ((rc = load_S())) return rc;
    if ((rc = factor_and_solve(c, load_S, load_h))) return rc;
    if (h_H_out) {
        CK(cudaEventRecord(c->ev_H_done, c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, c->ev_H_done, 0));
        c->out_rc = 0;
        c->out_thread = std::thread([c, h_H_out, n]() {
            cudaError_t e = cudaSetDevice(c->device);
            if (e == cudaSuccess) e = cudaMemcpy2DAsync(h_H_out, n * 8, c->H, (size_t)c->ld * 8, n * 8, n, cudaMemcpyDeviceToHost, c->copy_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->copy_stream);
            if (e != cudaSuccess) { c->out_rc = DYB_ECUDA; c->out_err = std::string("download of H' failed: ") + cudaGetErrorString(e); }
        });
    }
    return DYB_OK;
}

int dyb_form_hprime(dyb_ctx* c, const double* h_S, const double* h_h, double* h_H_out) {
    int rc = dyb_form_hprime_async(c, h_S, h_h, h_H_out);
    if (rc) return rc;
    if ((rc = dyb_wait_outputs(c))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

// ---- distributed formation on a team (team.cu): the Cholesky factor of S is computed once (first device), every member
// solves S X = h for ITS block of columns (2 N^3 / P flops each, in parallel) and the column blocks are then exchanged
// into row blocks by peer copies.  2.33 N^3 flops on one GPU become N^3/3 + 2 N^3/P.
// Factorisation only: S (host) -> Cholesky factor in ctx->S.  Full-matrix contexts.  DYB_ESINGULAR if S is not
// numerically positive definite (the caller then takes the single-GPU route with its LU fallback).
int dyb_factor_overlap(dyb_ctx* c, const double* h_S) {
    if (!c || !h_S) return fail(DYB_EINVAL, "NULL argument");
    CK(cudaSetDevice(c->device));
    int rc = dyb_wait_outputs(c);
    if (rc) return rc;
    if ((rc = ensure_solver(c))) return rc;
    const int64_t n = c->N;
    c->have_factor = false;
    CK(cudaMemcpyAsync(c->S, h_S, (size_t)n * n * 8, cudaMemcpyHostToDevice, c->stream));
    size_t wd = 0, wh = 0;
    int* d_info = reinterpret_cast<int*>(c->scal + 32);
    CKS(cusolverDnXpotrf_bufferSize(c->solver, c->sparams, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, c->S, n, CUDA_R_64F, &wd, &wh));
    DevScratch w1; std::vector<char> h_work(wh ? wh : 1);
    if (wd) CK(cudaMalloc(&w1.p, wd));
    CKS(cusolverDnXpotrf(c->solver, c->sparams, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, c->S, n, CUDA_R_64F, w1.p, wd, h_work.data(), wh, d_info));
    CK(cudaMemcpyAsync(c->h_scal + 100, d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (*reinterpret_cast<int*>(c->h_scal + 100) != 0) return fail(DYB_ESINGULAR, "S is not numerically positive definite (potrf info=%d)", *reinterpret_cast<int*>(c->h_scal + 100));
    c->have_factor = true; c->factor_is_lu = false;
    return DYB_OK;
}

int dyb_factor_device(dyb_ctx* c, void** d_U, int64_t* ldu) {
    if (!c || !d_U || !ldu) return fail(DYB_EINVAL, "NULL argument");
    if (!c->have_factor || c->factor_is_lu) return fail(DYB_EINVAL, "no Cholesky factor in this context");
    *d_U = c->S; *ldu = c->N;
    return DYB_OK;
}

// Member of a row-sharded team: bring columns row0 .. row0+M-1 of the host matrix h into the member's column block.
int dyb_upload_column_block(dyb_ctx* c, const double* h_h) {
    if (!c || !h_h) return fail(DYB_EINVAL, "NULL argument");
    CK(cudaSetDevice(c->device));
    int rc = dyb_wait_outputs(c);
    if (rc) return rc;
    if (!c->colblk) CK(cudaMalloc(&c->colblk, (size_t)c->N * c->M * 8));
    CK(cudaMemcpyAsync(c->colblk, h_h + (size_t)c->row0 * c->N, (size_t)c->N * c->M * 8, cudaMemcpyHostToDevice, c->stream));   // contiguous in column-major
    return DYB_OK;
}

// X = S^-1 h[:, block] with the Cholesky factor d_U (possibly on another device: copied over NVLink first); the solved
// block (= columns row0.. of H') is sent to h_H_out[:, block] beside whatever comes next (dyb_wait_outputs joins it).
int dyb_solve_column_block(dyb_ctx* c, const void* d_U, int64_t ldu, double* h_H_out) {
    if (!c || !d_U || ldu < c->N) return fail(DYB_EINVAL, "bad argument");
    if (!c->colblk) return fail(DYB_EINVAL, "dyb_upload_column_block must come first");
    CK(cudaSetDevice(c->device));
    int rc = ensure_solver(c);
    if (rc) return rc;
    const int64_t n = c->N;
    if (d_U != c->S) CK(cudaMemcpy2DAsync(c->S, (size_t)n * 8, d_U, (size_t)ldu * 8, (size_t)n * 8, n, cudaMemcpyDefault, c->stream));
    int* d_info = reinterpret_cast<int*>(c->scal + 32);
    CKS(cusolverDnXpotrs(c->solver, c->sparams, CUBLAS_FILL_MODE_UPPER, n, c->M, CUDA_R_64F, c->S, n, CUDA_R_64F, c->colblk, n, d_info));
    if (h_H_out) {
        CK(cudaEventRecord(c->ev_H_done, c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, c->ev_H_done, 0));
        c->out_rc = 0;
        double* dst = h_H_out + (size_t)c->row0 * n;
        const size_t bytes = (size_t)n * c->M * 8;
        c->out_thread = std::thread([c, dst, bytes]() {
            cudaError_t e = cudaSetDevice(c->device);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dst, c->colblk, bytes, cudaMemcpyDeviceToHost, c->copy_stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->copy_stream);
            if (e != cudaSuccess) { c->out_rc = DYB_ECUDA; c->out_err = std::string("download of H' failed: ") + cudaGetErrorString(e); }
        });
    }
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

int dyb_column_block_device(dyb_ctx* c, void** d_X) {
    if (!c || !d_X) return fail(DYB_EINVAL, "NULL argument");
    *d_X = c->colblk;
    return DYB_OK;
}

// The member's row block of H' out of the P solved column blocks (d_X[q] = columns q*M .. of H', N rows, ld N): P peer
// copies of M x M sub-blocks over NVLink (the block-transposed all-to-all of the distributed formation).
int dyb_take_rows_from_column_blocks(dyb_ctx* c, void* const* d_X, int n_blocks) {
    if (!c || !d_X || n_blocks != c->world) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    const size_t n = c->N, M = c->M;
    for (int q = 0; q < n_blocks; ++q) {
        if (!d_X[q]) return fail(DYB_EINVAL, "column block %d missing", q);
        CK(cudaMemcpy2DAsync(c->H + (size_t)q * M * c->ld, (size_t)c->ld * 8, static_cast<const double*>(d_X[q]) + c->row0, n * 8,
                             M * 8, M, cudaMemcpyDefault, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    c->have_factor = false;              // members keep a copy of the factor only as scratch; S^-1 applications go through the first device
    return DYB_OK;
}

// Build_Huckel on the device (ElHl_Chebyshev.f:296-323, X_ij of hamiltonians.f:33-63): h(i,j) = X_ij * S(i,j), computed
// from the upper triangle of S and mirrored like the reference's loop (j = 1..N, i = 1..j).  With this only S and
// three per-orbital vectors cross PCIe instead of S and h.
__global__ void build_huckel_kernel(int n, const double* __restrict__ S, const double* __restrict__ IP,
                                    const double* __restrict__ kWH, const double* __restrict__ Vs, double* __restrict__ h, long long ldh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= n) return;
    const int a = min(i, j), b = max(i, j);
    const double s = S[(size_t)a + (size_t)b * n];
    double x;
    if (i == j) x = IP[i] + Vs[i];
    else {
        const double c1 = IP[a] - IP[b], c2 = IP[a] + IP[b];
        const double c3 = (c1 / c2) * (c1 / c2);
        const double c4 = (Vs[a] + Vs[b]) * 0.5;
        const double kw = (kWH[a] + kWH[b]) * 0.5;
        const double keff = kw + c3 + c3 * c3 * (1.0 - kw);
        x = keff * c2 * 0.5 + c4;
    }
    h[(size_t)i + (size_t)j * ldh] = x * s;
}

int dyb_form_hprime_from_overlap(dyb_ctx* c, const double* h_S, const double* IP, const double* k_WH, const double* V_shift, double* h_H_out) {
    if (!c || !h_S || !IP || !k_WH || !V_shift) return fail(DYB_EINVAL, "NULL argument");
    if (c->M != c->N) return fail(DYB_EINVAL, "full-matrix contexts only");
    CK(cudaSetDevice(c->device));
    int rc = ensure_solver(c);
    if (rc) return rc;
    const size_t n = c->N;
    auto load_S = [&]() -> int { CK(cudaMemcpyAsync(c->S, h_S, n * n * 8, cudaMemcpyHostToDevice, c->stream)); return DYB_OK; };
    if ((rc = load_S())) return rc;
    double* par = c->io;                                            // 3 n doubles of the 8 n staging buffer
    CK(cudaMemcpyAsync(par, IP, n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(par + n, k_WH, n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(par + 2 * n, V_shift, n * 8, cudaMemcpyHostToDevice, c->stream));
    build_huckel_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, c->stream>>>((int)n, c->S, par, par + n, par + 2 * n, c->H, c->ld);
    c->launches++;
    CK(cudaGetLastError());
    if ((rc = factor_and_solve(c, load_S))) return rc;
    if (h_H_out) CK(cudaMemcpy2DAsync(h_H_out, n * 8, c->H, (size_t)c->ld * 8, n * 8, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

int dyb_form_hprime_device(dyb_ctx* c, const void* d_S, int64_t lds, const void* d_h, int64_t ldh) {
    if (!c || !d_S || !d_h || lds < c->N || ldh < c->N) return fail(DYB_EINVAL, "bad argument");
    if (c->M != c->N) return fail(DYB_EINVAL, "dyb_form_hprime_device needs the full matrix on one device");
    CK(cudaSetDevice(c->device));
    int rc = ensure_solver(c);
    if (rc) return rc;
    const size_t n = c->N;
    auto load_S = [&]() -> int { CK(cudaMemcpy2DAsync(c->S, n * 8, d_S, (size_t)lds * 8, n * 8, n, cudaMemcpyDeviceToDevice, c->stream)); return DYB_OK; };
    if ((rc = load_S())) return rc;
    CK(cudaMemcpy2DAsync(c->H, (size_t)c->ld * 8, d_h, (size_t)ldh * 8, n * 8, n, cudaMemcpyDeviceToDevice, c->stream));
    if ((rc = factor_and_solve(c, load_S))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

// ---- packets --------------------------------------------------------------------------------------
static int upload_quad(dyb_ctx* c, int n_part, const dyb_complex* src, double* dst_quad /* indexed by global index */) {
    const size_t n = c->N;
    double2* stage = reinterpret_cast<double2*>(c->io);
    CK(cudaMemcpyAsync(stage, src, n * n_part * sizeof(dyb_complex), cudaMemcpyHostToDevice, c->stream));
    pack_quad_kernel<<<(2 * c->N + 255) / 256, 256, 0, c->stream>>>(c->N, n_part, stage, dst_quad);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}
static int download_quad(dyb_ctx* c, int n_part, const double* src_quad, dyb_complex* dst) {
    const size_t n = c->N;
    double2* stage = reinterpret_cast<double2*>(c->io);
    unpack_quad_kernel<<<(2 * c->N + 255) / 256, 256, 0, c->stream>>>(c->N, n_part, src_quad, stage);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dst, stage, n * n_part * sizeof(dyb_complex), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

// Full host packets in, on every rank.  Ket-type vectors live at their global index on every rank; bra-type
// vectors are stored by LOCAL row (the dual product consumes x_bra at the shard's rows only).
int dyb_set_packets(dyb_ctx* c, int n_part, const dyb_complex* bra, const dyb_complex* ket) {
    if (!c || !bra || !ket || n_part < 1 || n_part > 2) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    int rc;
    if (c->M == c->N) {
        if ((rc = upload_quad(c, n_part, bra, c->psi_b))) return rc;
    } else {
        if (!c->full_tmp) return fail(DYB_EINVAL, "row-sharded context: call dyb_comm_init first");
        if ((rc = upload_quad(c, n_part, bra, c->full_tmp))) return rc;
        CK(cudaMemcpyAsync(c->psi_b, c->full_tmp + (size_t)c->row0 * NQ, (size_t)c->M * NQ * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    if ((rc = upload_quad(c, n_part, ket, c->psi_k))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    c->n_part = n_part;
    return DYB_OK;
}

// Full host packets out, on every rank (row-sharded: the owned slices are all-gathered first; collective call).
int dyb_get_packets(dyb_ctx* c, int n_part, dyb_complex* bra, dyb_complex* ket) {
    if (!c || !bra || !ket || n_part < 1 || n_part > 2) return fail(DYB_EINVAL, "bad argument");
    CK(cudaSetDevice(c->device));
    int rc;
    if (c->M == c->N) {
        if ((rc = download_quad(c, n_part, c->psi_b, bra))) return rc;
        if ((rc = download_quad(c, n_part, c->psi_k, ket))) return rc;
        return DYB_OK;
    }
    if (!c->comm) return fail(DYB_EINVAL, "row-sharded context: call dyb_comm_init first");
    CKN(g_nccl.AllGather(c->psi_b, c->full_tmp, (size_t)c->M * NQ, ncclDouble, c->comm, c->stream));
    if ((rc = download_quad(c, n_part, c->full_tmp, bra))) return rc;
    CKN(g_nccl.AllGather(c->psi_k + (size_t)c->row0 * NQ, c->full_tmp, (size_t)c->M * NQ, ncclDouble, c->comm, c->stream));
    if ((rc = download_quad(c, n_part, c->full_tmp, ket))) return rc;
    return DYB_OK;
}

// ---- row-sharded operation: NCCL communicator (one process per GPU; the unique id travels by the host layer)
int dyb_comm_unique_id(char* out128) {
    if (!out128) return fail(DYB_EINVAL, "NULL argument");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    CKN(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return DYB_OK;
}

// ---- fused peer-memory exchange: every rank exports one IPC buffer, then maps the peers' buffers --------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// the exchange buffer of the fused peer-memory path (allocated once; the rotating ket vectors move into it)
static int p2p_alloc(dyb_ctx* c) {
    if (!c->comm || c->world < 2) return fail(DYB_EINVAL, "dyb_comm_init (world >= 2) must come first");
    if (c->world > MAX_PEERS) return fail(DYB_EINVAL, "at most %d ranks", MAX_PEERS);
    CK(cudaSetDevice(c->device));
    if (const char* e = getenv("DYNEMOL_B200_PEER_TIMEOUT_S")) {
        const double sec = atof(e);
        if (sec > 0.0) { const unsigned long long ns = (unsigned long long)(sec * 1e9); CK(cudaMemcpyToSymbol(g_peer_timeout_ns, &ns, sizeof ns)); }
    }
    if (!c->comm_buf) {
        const size_t vec = align_up(c->Lq * NQ * sizeof(double), 4096);
        size_t off = 0;
        for (int i = 0; i < 2; ++i) { c->off_rs[i] = off; off += vec; }
        for (int i = 0; i < 3; ++i) { c->off_vk[i] = off; off += vec; }
        for (int i = 0; i < 2; ++i) { c->off_scal[i] = off; off += align_up((size_t)c->world * 8 * sizeof(double), 4096); }
        c->off_ready = off; off += 4096;
        c->off_done = off;  off += 4096;
        c->comm_bytes = off;
        CK(cudaMalloc(&c->comm_buf, off));
        CK(cudaMemset(c->comm_buf, 0, off));
        CK(cudaDeviceSynchronize());
        // the rotating ket vectors move into the shared buffer: the peers store their slices straight into them
        CK(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < 3; ++i) { if (c->vk[i]) cudaFree(c->vk[i]); c->vk[i] = reinterpret_cast<double*>(c->comm_buf + c->off_vk[i]); }
        c->vk_in_comm = true;
    }
    return DYB_OK;
}

int dyb_comm_p2p_handle(dyb_ctx* c, char* out64) {
    if (!c || !out64) return fail(DYB_EINVAL, "NULL argument");
    int rc = p2p_alloc(c);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->comm_buf));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out64, &h, 64);
    return DYB_OK;
}

int dyb_comm_p2p_open(dyb_ctx* c, const char* handles /* world x 64 bytes, rank order */) {
    if (!c || !handles) return fail(DYB_EINVAL, "NULL argument");
    if (!c->comm_buf) return fail(DYB_EINVAL, "dyb_comm_p2p_handle must come first");
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->peer_base[r] = c->comm_buf; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_base[r] = static_cast<char*>(ptr);
    }
    c->p2p = true;
    return DYB_OK;
}

// Same-process peers (dyb_team: one host thread per GPU): no IPC, the peers' buffers are mapped by enabling peer access
// between the devices.  Two phases so that every member has allocated its buffer before anybody maps it:
// phase 0 allocates, phase 1 (after a barrier of the caller) maps.  members: world contexts in rank order.
int dyb_comm_p2p_open_local(dyb_ctx* c, dyb_ctx* const* members, int phase) {
    if (!c || !members) return fail(DYB_EINVAL, "NULL argument");
    if (phase == 0) return p2p_alloc(c);
    if (!c->comm_buf) return fail(DYB_EINVAL, "phase 0 must come first");
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (!members[r] || !members[r]->comm_buf) return fail(DYB_EINVAL, "member %d has no exchange buffer", r);
        if (r != c->rank) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->device, members[r]->device));
            if (!can) return fail(DYB_ECUDA, "device %d cannot access device %d", c->device, members[r]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(members[r]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(DYB_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        c->peer_base[r] = members[r]->comm_buf;
    }
    c->p2p = true; c->p2p_local = true;
    return DYB_OK;
}

// switch between the fused peer-memory exchange and the NCCL collectives (collective decision: every rank must pass
// the same value); enabling needs a successful dyb_comm_p2p_open
int dyb_comm_p2p_enable(dyb_ctx* c, int on) {
    if (!c) return fail(DYB_EINVAL, "ctx is NULL");
    if (on) {
        for (int r = 0; r < c->world; ++r) if (!c->peer_base[r]) return fail(DYB_EINVAL, "peer %d is not mapped: call dyb_comm_p2p_open first", r);
    }
    c->p2p = on != 0;
    return DYB_OK;
}

int dyb_comm_init(dyb_ctx* c, int rank, int world, const char* id128) {
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) return fail(DYB_EINVAL, "bad argument");
    if (c->N % world != 0 || c->M != c->N / world || c->row0 != rank * c->M)
        return fail(DYB_EINVAL, "row sharding must be uniform: N=%d world=%d rank=%d row0=%d n_rows=%d", c->N, world, rank, c->row0, c->M);
    int rc = nccl_load();
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    CKN(g_nccl.CommInitRank(&c->comm, world, id, rank));
    c->rank = rank; c->world = world;
    if ((rc = alloc_zero(&c->rs_send, (size_t)c->Lq * NQ))) return rc;
    if ((rc = alloc_zero(&c->rs_recv, (size_t)c->Lq * NQ))) return rc;
    if ((rc = alloc_zero(&c->scal_all, (size_t)world * 8))) return rc;
    if ((rc = alloc_zero(&c->full_tmp, (size_t)c->Lq * NQ))) return rc;
    CK(cudaDeviceSynchronize());
    return DYB_OK;
}

// ---- propagation ----------------------------------------------------------------------------------
int dyb_propagate(dyb_ctx* c, int mode, double t_init, double t_max, const double* tau, double* save_tau, dyb_trace* traces) {
    if (!c || !tau || !save_tau) return fail(DYB_EINVAL, "NULL argument");
    if (c->n_part < 1) return fail(DYB_EINVAL, "dyb_set_packets must be called first");
    CK(cudaSetDevice(c->device));
    if ((mode == DYB_MODE_TAYLOR_REFGPU || mode == DYB_MODE_CHEBYSHEV_REFGPU) && c->world > 1)
        return fail(DYB_EINVAL, "the reference-GPU parity modes are single-GPU (like the reference's GPU path)");
    if (mode == DYB_MODE_TAYLOR || mode == DYB_MODE_TAYLOR_REFGPU || mode == DYB_MODE_CHEBYSHEV_REFGPU)
        return propagate_series(c, mode, t_init, t_max, tau, save_tau, traces);
    if (mode == DYB_MODE_CHEBYSHEV || mode == DYB_MODE_CHEBYSHEV_FULL) {
        if (!c->have_bounds) return fail(DYB_EINVAL, "Chebyshev mode needs spectral bounds: dyb_set_spectral_bounds / dyb_estimate_spectral_bounds");
        if (mode == DYB_MODE_CHEBYSHEV_FULL) return propagate_cheb_full(c, t_init, t_max, tau, save_tau, traces);
        return propagate_series(c, mode, t_init, t_max, tau, save_tau, traces);
    }
    return fail(DYB_EINVAL, "unknown mode %d", mode);
}

int dyb_run_terms(dyb_ctx* c, double tau, int n_terms, float* elapsed_ms, float* kernel_ms) {
    if (!c || n_terms < 1) return fail(DYB_EINVAL, "bad argument");
    if (c->n_part < 1) return fail(DYB_EINVAL, "dyb_set_packets must be called first");
    CK(cudaSetDevice(c->device));
    cplx C[ORDER];
    taylor_coefficient(tau, C);
    const bool per_kernel = kernel_ms != nullptr;
    const size_t need = 2 + (per_kernel ? 2 * (size_t)n_terms : 0);
    while (c->ev.size() < need) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->ev.push_back(e); }
    const int none[2] = {0, 0}, both[2] = {1, c->n_part > 1 ? 1 : 0};
    int rc, cur = 0, nxt = 1;
    if (single_launch_ok(c) && !per_kernel) {
        // one series_init + one cooperative launch per 24-term series
        CK(cudaEventRecord(c->ev[0], c->stream));
        for (int s0 = 0; s0 < n_terms; s0 += ORDER - 1) {
            const int len = std::min(ORDER - 1, n_terms - s0);
            if ((rc = launch_series_init(c, none, both, 0))) return rc;
            std::vector<PassParams> passes(len);
            for (int s = 0; s < len; ++s) {
                memset(&passes[s], 0, sizeof(PassParams));
                const int k = 2 + s;
                const cplx r = C[k - 1] / C[k - 2];
                for (int p = 0; p < 2; ++p) {
                    PartPass& a = passes[s].part[p];
                    a.active = both[p]; a.k = k; a.alpha_re = r.real(); a.alpha_im = r.imag(); a.norm_ref = 1.0;
                }
            }
            if ((rc = run_series_single_launch(c, passes))) return rc;
        }
        CK(cudaEventRecord(c->ev[1], c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->passes_last = n_terms;
        if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, c->ev[0], c->ev[1]));
        return DYB_OK;
    }
    CK(cudaEventRecord(c->ev[0], c->stream));
    for (int s = 0; s < n_terms; ++s) {
        const int k = 2 + (s % (ORDER - 1));                      // repeated 24-term series, like Convergence calls
        if (k == 2) { cur = 0; nxt = 1; if ((rc = launch_series_init(c, none, both, cur))) return rc; }
        EpiParams E = epi_params(c, cur, cur, nxt);
        const cplx r = C[k - 1] / C[k - 2];
        for (int p = 0; p < 2; ++p) {
            PartPass& a = E.pass.part[p];
            a.active = both[p]; a.k = k; a.alpha_re = r.real(); a.alpha_im = r.imag(); a.norm_ref = 1.0;
        }
        if (per_kernel) {
            CK(cudaEventRecord(c->ev[2 + 2 * s], c->stream));
            if ((rc = launch_matvec(c, c->vk[cur], c->vb[cur], false))) return rc;
            CK(cudaEventRecord(c->ev[3 + 2 * s], c->stream));
            if (c->world > 1) return fail(DYB_EINVAL, "per-kernel timing is a single-GPU diagnostic");
            if ((rc = launch_epilogue(c, E))) return rc;
        } else if ((rc = run_term(c, E, cur, nxt, false))) return rc;
        std::swap(cur, nxt);
    }
    CK(cudaEventRecord(c->ev[1], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->world > 1 && (rc = read_ctrl(c))) return rc;             // a peer-flag timeout surfaces here
    c->passes_last = n_terms;
    if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, c->ev[0], c->ev[1]));
    if (per_kernel) {
        float tot = 0.f;
        for (int s = 0; s < n_terms; ++s) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev[2 + 2 * s], c->ev[3 + 2 * s])); tot += ms; }
        *kernel_ms = tot;
    }
    return DYB_OK;
}

int dyb_dual_matvec(dyb_ctx* c, int n_part, const dyb_complex* xb, const dyb_complex* xk, dyb_complex* yb, dyb_complex* yk) {
    if (!c || !xb || !xk || !yb || !yk || n_part < 1 || n_part > 2) return fail(DYB_EINVAL, "bad argument");
    if (c->M != c->N) return fail(DYB_EINVAL, "full-matrix contexts only");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = upload_quad(c, n_part, xb, c->vb[0]))) return rc;
    if ((rc = upload_quad(c, n_part, xk, c->vk[0]))) return rc;
    if ((rc = launch_matvec(c, c->vk[0], c->vb[0], false))) return rc;
    EpiParams E = epi_params(c, 0, 0, 1);
    slab_reduce_kernel<<<(2 * c->M + 255) / 256, 256, 0, c->stream>>>(E);
    c->launches++;
    CK(cudaGetLastError());
    if ((rc = download_quad(c, n_part, c->vb[1], yb))) return rc;
    if ((rc = download_quad(c, n_part, c->vk[1], yk))) return rc;
    return DYB_OK;
}

// ---- spectral bounds for the Chebyshev mode ---------------------------------------------------------
int dyb_set_spectral_bounds(dyb_ctx* c, double emin, double emax) {
    if (!c || !(emax > emin)) return fail(DYB_EINVAL, "need emax > emin");
    c->emin = emin; c->emax = emax; c->have_bounds = true;
    return DYB_OK;
}

int dyb_get_spectral_bounds(dyb_ctx* c, double* emin, double* emax) {
    if (!c || !emin || !emax) return fail(DYB_EINVAL, "NULL argument");
    if (!c->have_bounds) return fail(DYB_EINVAL, "no spectral bounds set");
    *emin = c->emin; *emax = c->emax;
    return DYB_OK;
}

// number of eigenvalues of the symmetric tridiagonal (a, b) below x (Sturm sequence)
static int sturm_count(const std::vector<double>& a, const std::vector<double>& b, double x) {
    int cnt = 0; double d = 1.0;
    for (size_t i = 0; i < a.size(); ++i) {
        d = a[i] - x - (i ? b[i] * b[i] / d : 0.0);
        if (d == 0.0) d = 1e-300;
        if (d < 0.0) ++cnt;
    }
    return cnt;
}
static double tridiag_eig(const std::vector<double>& a, const std::vector<double>& b, int which /*0 = smallest, else largest*/) {
    double lo = 1e300, hi = -1e300;
    for (size_t i = 0; i < a.size(); ++i) {
        const double r = (i ? fabs(b[i]) : 0.0) + (i + 1 < a.size() ? fabs(b[i + 1]) : 0.0);
        lo = std::min(lo, a[i] - r); hi = std::max(hi, a[i] + r);
    }
    const int target = which ? (int)a.size() - 1 : 0;          // index of the wanted eigenvalue
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (sturm_count(a, b, mid) > target) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

// In-place sum over the ranks of a row-sharded run (no-op on one GPU)
static int allreduce_sum(dyb_ctx* c, double* buf, size_t n) {
    if (c->world > 1) CKN(g_nccl.AllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, c->stream));
    return DYB_OK;
}

// One dual product with no series state: yk[row0 + i] = (H' xk)_i for the owned rows, yb[i] = (H'^T xb)_(row0+i).
// xk: full-length ket vector (global index, zero padded), xb: owned slice of the bra vector (zero padded to the panels).
static int plain_dual_matvec(dyb_ctx* c, const double* xk, const double* xb, double* yk, double* yb) {
    int rc;
    if ((rc = launch_matvec(c, xk, xb, false))) return rc;
    if (c->world > 1) {
        const int n2 = 2 * c->N;
        bra_panel_reduce_kernel<<<(n2 + 255) / 256, 256, 0, c->stream>>>(c->N, c->NP, c->Ncpad, c->bra_slab, c->rs_send);
        c->launches++;
        CK(cudaGetLastError());
        CKN(g_nccl.ReduceScatter(c->rs_send, c->rs_recv, (size_t)c->M * NQ, ncclDouble, ncclSum, c->comm, c->stream));
    }
    EpiParams E = epi_params(c, 0, 0, 1);
    E.nxt_k = yk; E.nxt_b = yb;
    slab_reduce_kernel<<<(2 * c->M + 255) / 256, 256, 0, c->stream>>>(E);
    c->launches++;
    CK(cudaGetLastError());
    return DYB_OK;
}

// The owned ket slices of all ranks -> one full-length vector on every rank (one GPU: a copy)
static int gather_ket(dyb_ctx* c, const double* slice, double* full) {
    if (c->world > 1) CKN(g_nccl.AllGather(slice, full, (size_t)c->M * NQ, ncclDouble, c->comm, c->stream));
    else CK(cudaMemcpyAsync(full, slice, (size_t)c->M * NQ * 8, cudaMemcpyDeviceToDevice, c->stream));
    return DYB_OK;
}

// Lanczos in the S inner product started from the current packets (w0 = Psi_bra = S v0, v0 = Psi_ket): because
// H'^T S = S H', the left vectors stay w_j = S v_j and the recurrence is the symmetric one; every step is one
// dual product (H' v_j and H'^T w_j) of the hot kernel.  Ritz values lie inside the spectrum, hence `margin`
// (fraction of the width added on both sides).  Vectors, dot products and the scalar decisions live on the device
// (lanczos.cuh); the host downloads the tridiagonal coefficients once at the end.  Row-sharded contexts take part with
// their owned rows (collective call: partial dots are all-reduced, the ket vector is all-gathered before each product).
int dyb_estimate_spectral_bounds(dyb_ctx* c, int n_iter, double margin, double* emin_out, double* emax_out) {
    if (!c || n_iter < 2 || n_iter > LZ_MAX_IT) return fail(DYB_EINVAL, "bad argument (2 <= n_iter <= %d)", LZ_MAX_IT);
    if (c->n_part < 1) return fail(DYB_EINVAL, "dyb_set_packets must be called first (the packets start the Lanczos run)");
    if (c->world > 1 && !c->comm) return fail(DYB_EINVAL, "row-sharded context: call dyb_comm_init first");
    CK(cudaSetDevice(c->device));
    const int np = c->n_part, M = c->M;
    const size_t stride = (size_t)M * NQ;
    if (c->lz_cap < n_iter) {
        for (double** b : {&c->lz_V, &c->lz_W}) { if (*b) cudaFree(*b); *b = nullptr; }
        CK(cudaMalloc(&c->lz_V, (size_t)(n_iter + 1) * stride * 8));
        CK(cudaMalloc(&c->lz_W, (size_t)(n_iter + 1) * stride * 8));
        c->lz_cap = n_iter;
    }
    if (!c->lz_dots) CK(cudaMalloc(&c->lz_dots, (size_t)(2 * (LZ_MAX_IT + 1) * 4 + 16) * 8));
    if (!c->lz_state) CK(cudaMalloc(&c->lz_state, sizeof(LanczosState)));
    CK(cudaMemsetAsync(c->lz_state, 0, sizeof(LanczosState), c->stream));
    double* const dots = c->lz_dots;                                   // [2 (j+1)][2](re,im): Gram-Schmidt coefficients / single dots
    double* const coef = c->lz_dots + 2 * (LZ_MAX_IT + 1) * 4;         // three-term coefficient list (2 entries)
    double* const scale = coef + 8;
    const unsigned vg = (unsigned)((2 * M + 255) / 256);
    int rc;
    auto dots_of = [&](const double* X, int nq, const double* y, double* out) -> int {
        lz_dots_kernel<<<nq, LZ_THREADS, 0, c->stream>>>(M, X, stride, y, out);
        c->launches++;
        CK(cudaGetLastError());
        return DYB_OK;
    };
    auto combine = [&](const double* x, const double* X, int nq, const double* cf, const double* sc, double* y) -> int {
        lz_combine_kernel<<<vg, 256, 0, c->stream>>>(M, x, X, stride, nq, cf, sc, y);
        c->launches++;
        CK(cudaGetLastError());
        return DYB_OK;
    };
    auto scalar = [&](int op, int j) -> int {
        lz_scalar_kernel<<<1, 32, 0, c->stream>>>(op, j, dots, c->lz_state, coef, scale);
        c->launches++;
        CK(cudaGetLastError());
        return DYB_OK;
    };
    auto V = [&](int j) { return c->lz_V + (size_t)j * stride; };
    auto W = [&](int j) { return c->lz_W + (size_t)j * stride; };
    const double* v0 = c->psi_k + (size_t)c->row0 * NQ;                // owned slices of the packets
    const double* w0 = c->psi_b;
    // normalise: <w0|v0> = 1
    if ((rc = dots_of(w0, 1, v0, dots)) || (rc = allreduce_sum(c, dots, 4)) || (rc = scalar(LZ_OP_START, 0))) return rc;
    if ((rc = combine(v0, v0, 0, dots, scale, V(0))) || (rc = combine(w0, w0, 0, dots, scale, W(0)))) return rc;
    for (int j = 0; j < n_iter; ++j) {
        // H' v_j and H'^T w_j: one dual product
        if ((rc = gather_ket(c, V(j), c->vk[0]))) return rc;
        CK(cudaMemcpyAsync(c->vb[0], W(j), stride * 8, cudaMemcpyDeviceToDevice, c->stream));
        if ((rc = plain_dual_matvec(c, c->vk[0], c->vb[0], c->vk[1], c->vb[1]))) return rc;
        const double* hv = c->vk[1] + (size_t)c->row0 * NQ;
        const double* hw = c->vb[1];
        // alpha_j = <w_j|H'v_j> ; v' = H'v_j - alpha_j v_j - beta_j v_{j-1} (same for w')
        if ((rc = dots_of(W(j), 1, hv, dots)) || (rc = allreduce_sum(c, dots, 4)) || (rc = scalar(LZ_OP_ALPHA, j))) return rc;
        const int j0 = j > 0 ? j - 1 : 0, n3 = j > 0 ? 2 : 1;
        if ((rc = combine(hv, V(j0), n3, coef, nullptr, V(j + 1))) || (rc = combine(hw, W(j0), n3, coef, nullptr, W(j + 1)))) return rc;
        // full two-sided Gram-Schmidt against every previous pair, applied twice: w_q^H v' = 0, v_q^H w' = 0
        for (int sweep = 0; sweep < 2; ++sweep) {
            double* dA = dots; double* dB = dots + (size_t)(j + 1) * 4;
            if ((rc = dots_of(W(0), j + 1, V(j + 1), dA)) || (rc = dots_of(V(0), j + 1, W(j + 1), dB))) return rc;
            if ((rc = allreduce_sum(c, dots, (size_t)2 * (j + 1) * 4))) return rc;
            if ((rc = combine(V(j + 1), V(0), j + 1, dA, nullptr, V(j + 1))) || (rc = combine(W(j + 1), W(0), j + 1, dB, nullptr, W(j + 1)))) return rc;
        }
        // beta_{j+1}^2 = <w'|v'> ; breakdown test ; normalise
        if ((rc = dots_of(W(j + 1), 1, V(j + 1), dots)) || (rc = allreduce_sum(c, dots, 4)) || (rc = scalar(LZ_OP_BETA, j))) return rc;
        if ((rc = combine(V(j + 1), V(0), 0, dots, scale, V(j + 1))) || (rc = combine(W(j + 1), W(0), 0, dots, scale, W(j + 1)))) return rc;
    }
    static_assert(sizeof(LanczosState) < (1 << 16), "LanczosState");
    std::vector<char> hbuf(sizeof(LanczosState));
    CK(cudaMemcpyAsync(hbuf.data(), c->lz_state, sizeof(LanczosState), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const LanczosState& st = *reinterpret_cast<const LanczosState*>(hbuf.data());
    double lo = 1e300, hi = -1e300;
    for (int p = 0; p < np; ++p) {
        if (st.bad_start[p]) return fail(DYB_EINVAL, "<bra|ket> of particle %d is not positive: packets are not an S-dual pair", p);
        std::vector<double> al, be;
        for (int j = 0; j < n_iter; ++j) {            // a step is only accepted once its successor is sound
            if (st.ok[p][j]) { al.push_back(st.alpha[p][j]); be.push_back(st.beta[p][j]); }
            else { if (j == 0) { al.push_back(st.alpha[p][0]); be.push_back(0.0); } break; }
        }
        if (al.empty()) continue;
        lo = std::min(lo, tridiag_eig(al, be, 0)); hi = std::max(hi, tridiag_eig(al, be, 1));
    }
    if (!(hi > lo)) return fail(DYB_EINVAL, "Lanczos produced a degenerate interval");
    const double width = hi - lo;
    c->emin = lo - margin * width; c->emax = hi + margin * width; c->have_bounds = true;
    if (emin_out) *emin_out = c->emin;
    if (emax_out) *emax_out = c->emax;
    return DYB_OK;
}

// ---- post-step quantities -------------------------------------------------------------------------
// AO_bra = S^-1 Psi_bra (ElHl_Chebyshev.f:274, Taylor_gpu.cpp:718): the complex vector is solved as its
// real and imaginary parts against the kept factor of S.
int dyb_ao_bra(dyb_ctx* c, int n_part, dyb_complex* h_AO_bra) {
    if (!c || !h_AO_bra || n_part < 1 || n_part > 2) return fail(DYB_EINVAL, "bad argument");
    if (!c->have_factor) return fail(DYB_EINVAL, "dyb_ao_bra needs the factor of S: call dyb_form_hprime first");
    CK(cudaSetDevice(c->device));
    const int64_t n = c->N;
    // stage: complex columns (n x n_part) in c->io; view as real (2n x n_part)?  potrs needs separate real
    // right-hand sides of length n, so de-interleave on the host side of the staging buffer with cublasDcopy.
    double2* stage = reinterpret_cast<double2*>(c->io);
    unpack_quad_kernel<<<(2 * c->N + 255) / 256, 256, 0, c->stream>>>(c->N, n_part, c->psi_b, stage);
    c->launches++;
    CK(cudaGetLastError());
    if (!c->blas) { CKB(cublasCreate(&c->blas)); CKB(cublasSetStream(c->blas, c->stream)); }
    double* rhs = c->io + (size_t)n * 4;                  // 2*n_part real columns of length n
    for (int p = 0; p < n_part; ++p) {
        CKB(cublasDcopy(c->blas, (int)n, c->io + (size_t)p * 2 * n, 2, rhs + (size_t)(2 * p) * n, 1));
        CKB(cublasDcopy(c->blas, (int)n, c->io + (size_t)p * 2 * n + 1, 2, rhs + (size_t)(2 * p + 1) * n, 1));
    }
    int* d_info = reinterpret_cast<int*>(c->scal + 32);
    if (c->factor_is_lu) CKS(cusolverDnXgetrs(c->solver, c->sparams, CUBLAS_OP_N, n, 2 * n_part, CUDA_R_64F, c->S, n, c->ipiv, CUDA_R_64F, rhs, n, d_info));
    else CKS(cusolverDnXpotrs(c->solver, c->sparams, CUBLAS_FILL_MODE_UPPER, n, 2 * n_part, CUDA_R_64F, c->S, n, CUDA_R_64F, rhs, n, d_info));
    for (int p = 0; p < n_part; ++p) {
        CKB(cublasDcopy(c->blas, (int)n, rhs + (size_t)(2 * p) * n, 1, c->io + (size_t)p * 2 * n, 2));
        CKB(cublasDcopy(c->blas, (int)n, rhs + (size_t)(2 * p + 1) * n, 1, c->io + (size_t)p * 2 * n + 1, 2));
    }
    CK(cudaMemcpyAsync(h_AO_bra, c->io, (size_t)n * n_part * sizeof(dyb_complex), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return DYB_OK;
}

// QuasiParticleEnergies (ElHl_Chebyshev.f:329-371): erg(p) = sum_ij AO_bra(i,p) h(i,j) AO_ket(j,p) with
// AO_bra = conj(S^-1 Psi_bra), AO_ket = Psi_ket (ElHl_Chebyshev.f:274-276).  Because S^-1 is real symmetric,
//   erg = (S^-1 Psi_bra)^H h Psi_ket = Psi_bra^H (S^-1 h) Psi_ket = dotc(Psi_bra, H' Psi_ket):
// one more dual product of the hot kernel and a dot product, no second pass over h.  out = (re,im) per particle.
int dyb_quasiparticle_energies(dyb_ctx* c, int n_part, double* out_reim) {
    if (!c || !out_reim || n_part < 1 || n_part > 2) return fail(DYB_EINVAL, "bad argument");
    if (c->world > 1 && !c->comm) return fail(DYB_EINVAL, "row-sharded context: call dyb_comm_init first");
    CK(cudaSetDevice(c->device));
    int rc;
    const double* xk = c->psi_k;
    if (c->world > 1) {                    // row-sharded (collective call): every rank needs the whole ket; partial dots are summed
        if ((rc = gather_ket(c, c->psi_k + (size_t)c->row0 * NQ, c->vk[0]))) return rc;
        xk = c->vk[0];
    }
    if ((rc = plain_dual_matvec(c, xk, c->psi_b, c->vk[1], c->vb[1]))) return rc;      // vk[1] = H' Psi_ket on the owned rows
    dotc_kernel<<<1, 1024, 0, c->stream>>>(c->M, c->psi_b, c->vk[1] + (size_t)c->row0 * NQ, c->scal);
    c->launches++;
    CK(cudaGetLastError());
    if ((rc = allreduce_sum(c, c->scal, 4))) return rc;
    CK(cudaMemcpyAsync(c->h_scal, c->scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 2 * n_part; ++i) out_reim[i] = c->h_scal[i];
    return DYB_OK;
}

// Diabatic-Ehrenfest kernel (diabatic-Ehren.f:115-119; Taylor_gpu.cpp:743-797 ehrenfestkernel_gpu_):
//   K = X o A - H' A      (o = element-wise product; A = (rho + rho^T)/2, X = X_ij, both host-built)
// with the H' ALREADY RESIDENT from the propagation of this step: no 8 N^2 B upload of H' (the reference ships it
// to another MPI rank and uploads it again).  O(N^3) DGEMM through cuBLAS; the Hadamard-minus is fused in one pass.
__global__ void hadamard_minus_kernel(size_t n_elem, const double* __restrict__ X, const double* __restrict__ A, double* __restrict__ K) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_elem) K[i] = X[i] * A[i] - K[i];
}

// the three N x N scratch matrices (and the packet staging) of the Ehrenfest kernels live in the context: allocated on the
// first call, reused by every nuclear step (the reference allocates and frees them per call, Taylor_gpu.cpp:351-356 style)
static int ensure_ehrenfest_scratch(dyb_ctx* c) {
    const size_t bytes = (size_t)c->N * c->N * 8;
    if (!c->ehr_A) CK(cudaMalloc(&c->ehr_A, bytes));
    if (!c->ehr_X) CK(cudaMalloc(&c->ehr_X, bytes));
    if (!c->ehr_K) CK(cudaMalloc(&c->ehr_K, bytes));
    if (!c->ehr_vec) CK(cudaMalloc(&c->ehr_vec, (size_t)c->N * 2 * sizeof(dyb_complex) * 2));
    return DYB_OK;
}

int dyb_ehrenfest_kernel(dyb_ctx* c, const double* h_A, const double* h_X, double* h_K) {
    if (!c || !h_A || !h_X || !h_K) return fail(DYB_EINVAL, "NULL argument");
    if (c->M != c->N) return fail(DYB_EINVAL, "full-matrix contexts only");
    CK(cudaSetDevice(c->device));
    const size_t n = c->N, bytes = n * n * 8;
    { int rc0 = ensure_ehrenfest_scratch(c); if (rc0) return rc0; }
    double *A = c->ehr_A, *X = c->ehr_X, *K = c->ehr_K;
    int rc = DYB_OK;
    do {
        if (!c->blas) { if (cublasCreate(&c->blas) != CUBLAS_STATUS_SUCCESS || cublasSetStream(c->blas, c->stream) != CUBLAS_STATUS_SUCCESS) { rc = fail(DYB_ECUDA, "cublasCreate failed"); break; } }
        if (cudaMemcpyAsync(A, h_A, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(X, h_X, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = fail(DYB_ECUDA, "H2D of A/X failed"); break; }
        const double one = 1.0, zero = 0.0;
        if (cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, (int)n, (int)n, &one, c->H, (int)c->ld, A, (int)n, &zero, K, (int)n) != CUBLAS_STATUS_SUCCESS) { rc = fail(DYB_ECUDA, "cublasDgemm failed"); break; }
        hadamard_minus_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, c->stream>>>(n * n, X, A, K);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(h_K, K, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(DYB_ECUDA, "Ehrenfest kernel failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    return rc;
}

// The same kernel with the electron-hole density built on the device (Taylor_gpu.cpp:801-873 ehrenfestkernel2_gpu_,
// Chebyshev_gpu_kernels.cu:395-455 calculate_A; on the host: calculate_rho + A_ad_nd of diabatic-Ehren.f:107-109):
//   rho(i,j) = Re{ ket(j,1) bra(i,1) } - Re{ ket(j,2) bra(i,2) } ,  A = (rho + rho^T)/2 ,  K = X o A - H' A.
// bra/ket are the N x 2 complex AO packets (column 1 electron, column 2 hole); only 64 N bytes cross PCIe instead of 8 N^2.
__global__ void ehrenfest_density_kernel(int n, const double2* __restrict__ bra, const double2* __restrict__ ket, double* __restrict__ A) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * n) return;
    const int i = (int)(idx % n), j = (int)(idx / n);
    const double2 bi0 = bra[i], bi1 = bra[n + i], bj0 = bra[j], bj1 = bra[n + j];
    const double2 ki0 = ket[i], ki1 = ket[n + i], kj0 = ket[j], kj1 = ket[n + j];
    const double rho_ij = (kj0.x * bi0.x - kj0.y * bi0.y) - (kj1.x * bi1.x - kj1.y * bi1.y);
    const double rho_ji = (ki0.x * bj0.x - ki0.y * bj0.y) - (ki1.x * bj1.x - ki1.y * bj1.y);
    A[idx] = 0.5 * (rho_ij + rho_ji);
}

int dyb_ehrenfest_kernel2(dyb_ctx* c, const dyb_complex* h_bra, const dyb_complex* h_ket, const double* h_X, double* h_K) {
    if (!c || !h_bra || !h_ket || !h_X || !h_K) return fail(DYB_EINVAL, "NULL argument");
    if (c->M != c->N) return fail(DYB_EINVAL, "full-matrix contexts only");
    CK(cudaSetDevice(c->device));
    const size_t n = c->N, bytes = n * n * 8, vbytes = n * 2 * sizeof(dyb_complex);
    { int rc0 = ensure_ehrenfest_scratch(c); if (rc0) return rc0; }
    struct { void* p; } bB = {c->ehr_vec}, bKt = {reinterpret_cast<char*>(c->ehr_vec) + vbytes};
    double *A = c->ehr_A, *X = c->ehr_X, *K = c->ehr_K;
    int rc = DYB_OK;
    do {
        if (!c->blas) { if (cublasCreate(&c->blas) != CUBLAS_STATUS_SUCCESS || cublasSetStream(c->blas, c->stream) != CUBLAS_STATUS_SUCCESS) { rc = fail(DYB_ECUDA, "cublasCreate failed"); break; } }
        if (cudaMemcpyAsync(bB.p, h_bra, vbytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(bKt.p, h_ket, vbytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
            cudaMemcpyAsync(X, h_X, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = fail(DYB_ECUDA, "H2D of packets/X failed"); break; }
        ehrenfest_density_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, c->stream>>>((int)n, static_cast<const double2*>(bB.p), static_cast<const double2*>(bKt.p), A);
        c->launches++;
        const double one = 1.0, zero = 0.0;
        if (cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, (int)n, (int)n, &one, c->H, (int)c->ld, A, (int)n, &zero, K, (int)n) != CUBLAS_STATUS_SUCCESS) { rc = fail(DYB_ECUDA, "cublasDgemm failed"); break; }
        hadamard_minus_kernel<<<(unsigned)((n * n + 255) / 256), 256, 0, c->stream>>>(n * n, X, A, K);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(h_K, K, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = fail(DYB_ECUDA, "Ehrenfest kernel failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    return rc;
}

int dyb_populations(dyb_ctx* c, int n_part, int n_frag, const int32_t* fragment, double t, double* out) {
    if (!c || !fragment || !out || n_part < 1 || n_part > 2 || n_frag < 0 || n_frag > MAX_FRAG) return fail(DYB_EINVAL, "bad argument");
    if (c->world > 1 && !c->comm) return fail(DYB_EINVAL, "row-sharded context: call dyb_comm_init first");
    CK(cudaSetDevice(c->device));
    if (!c->frag) CK(cudaMalloc(&c->frag, sizeof(int) * c->N));
    CK(cudaMemcpyAsync(c->frag, fragment, sizeof(int) * c->N, cudaMemcpyHostToDevice, c->stream));
    double* d_out = c->io;                                  // 2*(MAX_FRAG+1) doubles
    if (n_part < 2) CK(cudaMemsetAsync(d_out, 0, sizeof(double) * 2 * (MAX_FRAG + 1), c->stream));
    // owned rows only; a row-sharded run (collective call) sums the per-rank partials
    populations_kernel<<<n_part, 256, 0, c->stream>>>(c->M, n_frag, c->frag + c->row0, c->psi_b, c->psi_k + (size_t)c->row0 * NQ, d_out);
    c->launches++;
    CK(cudaGetLastError());
    { int rc = allreduce_sum(c, d_out, 2 * (MAX_FRAG + 1)); if (rc) return rc; }
    CK(cudaMemcpyAsync(c->h_scal, d_out, sizeof(double) * 2 * (MAX_FRAG + 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int p = 0; p < n_part; ++p) {
        double* col = out + (size_t)p * (n_frag + 2);
        col[0] = t;
        for (int f = 0; f <= n_frag; ++f) col[1 + f] = c->h_scal[p * (MAX_FRAG + 1) + f];
    }
    return DYB_OK;
}

}  // extern "C"
