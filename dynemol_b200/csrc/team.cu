// team.cu -- single-process multi-GPU: a team of row-sharded contexts behind ONE host caller.
//
// the reference ABI is one host process (ElHl_Chebyshev_GPU.f:269-272 calls one C symbol), so the 8 GPUs of
// a box are reachable from Fortran only if the library drives them itself.  The row-sharded machinery of propagator.cu
// (one context per GPU: dual product on the owned rows, fused NVLink peer-memory exchange inside the epilogue kernel,
// replicated bit-identical decisions) was written for one process per GPU.  A team runs the very same per-context code
// from one host THREAD per device inside every call -- the contexts cannot tell the difference, except that their peers'
// exchange buffers are mapped by cudaDeviceEnablePeerAccess instead of CUDA IPC.  Built on the public dyb_* API only.
//
//   formation   H' = S^-1 h on the first device of the team (a full-size context: overlapped transfers, Cholesky, the
//               factor kept for AO_bra), then every member copies its row block over NVLink (peer D2D copy)
//   per term    dual product on the owned rows -> fused peer-memory exchange (reduce-scatter by peer loads, all-gather by
//               peer stores, epoch flags) -> replicated decision; no host thread talks to another one inside a series
//   elsewhere   NCCL (one communicator rank per thread) for the start-of-series all-gather, norm_ref, Lanczos dots,
//               populations, energies, packet gathers
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dynemol_b200.h"

struct dyb_team {
    int N = 0, P = 0;
    std::vector<int> dev;
    std::vector<dyb_ctx*> m;            // row-sharded members, rank order
    dyb_ctx* form = nullptr;            // full-size context on dev[0] for S^-1 h and S^-1 Psi_bra (created on first use)
    int n_part = 0;
    std::vector<std::vector<dyb_complex>> scratch;   // per-rank host packets (ranks > 0 of collective downloads)
};

namespace {

thread_local std::string t_err;
int tfail(int code, const std::string& msg) { t_err = msg; return code; }

// f(member, rank) on every member, rank 0 on the calling thread, the others on their own threads (the per-context code
// blocks in NCCL collectives and spins on peer flags exactly like separate processes would)
int team_run(dyb_team* t, const std::function<int(dyb_ctx*, int)>& f) {
    std::vector<int> rc(t->P, 0);
    std::vector<std::string> err(t->P);
    std::vector<std::thread> th;
    th.reserve(t->P);
    for (int r = 1; r < t->P; ++r)
        th.emplace_back([&, r]() { rc[r] = f(t->m[r], r); if (rc[r]) err[r] = dyb_last_error(); });
    rc[0] = f(t->m[0], 0);
    if (rc[0]) err[0] = dyb_last_error();
    for (auto& x : th) x.join();
    for (int r = 0; r < t->P; ++r)
        if (rc[r]) return tfail(rc[r], "team member " + std::to_string(r) + " (device " + std::to_string(t->dev[r]) + "): " + err[r]);
    return DYB_OK;
}

int ensure_form(dyb_team* t) {
    if (t->form) return DYB_OK;
    if (t->P == 1) { t->form = t->m[0]; return DYB_OK; }          // a team of one: the member is the full-size context
    int rc = dyb_create(&t->form, t->dev[0], t->N, 0, t->N);
    if (rc) return tfail(rc, std::string("formation context: ") + dyb_last_error());
    return DYB_OK;
}

}  // namespace

extern "C" {

const char* dyb_team_last_error(void) { return t_err.c_str(); }

int dyb_team_size(dyb_team* t) { return t ? t->P : 0; }

int dyb_team_destroy(dyb_team* t) {
    if (!t) return DYB_OK;
    // members first (their NCCL communicators are torn down together), then the formation context
    if (!t->m.empty()) {
        std::vector<std::thread> th;
        for (int r = 1; r < (int)t->m.size(); ++r) th.emplace_back([t, r]() { if (t->m[r]) dyb_destroy(t->m[r]); });
        if (t->m[0]) dyb_destroy(t->m[0]);
        for (auto& x : th) x.join();
    }
    if (t->form && t->P > 1) dyb_destroy(t->form);
    delete t;
    return DYB_OK;
}

int dyb_team_create(dyb_team** out, int n_dev, const int* devices, int N) {
    if (!out) return tfail(DYB_EINVAL, "out is NULL");
    *out = nullptr;
    const int have = dyb_device_count();
    if (have <= 0) return tfail(DYB_ENODEV, "no CUDA device available; dynemol_b200 has no CPU fallback");
    if (n_dev < 1 || n_dev > 8) return tfail(DYB_EINVAL, "a team has 1..8 devices");
    if (N <= 0 || N % n_dev != 0 || (N / n_dev) % 4 != 0)
        return tfail(DYB_EINVAL, "N = " + std::to_string(N) + " must split into " + std::to_string(n_dev) + " equal row blocks of a multiple of 4 rows");
    dyb_team* t = new dyb_team();
    t->N = N; t->P = n_dev;
    t->dev.resize(n_dev); t->m.assign(n_dev, nullptr); t->scratch.resize(n_dev);
    for (int r = 0; r < n_dev; ++r) {
        t->dev[r] = devices ? devices[r] : r;
        if (t->dev[r] < 0 || t->dev[r] >= have) { delete t; return tfail(DYB_EINVAL, "device " + std::to_string(devices ? devices[r] : r) + " out of range"); }
    }
    const int M = N / n_dev;
    char id[128];
    int rc = DYB_OK;
    if (n_dev > 1 && (rc = dyb_comm_unique_id(id))) { std::string e = dyb_last_error(); delete t; return tfail(rc, e); }
    rc = team_run(t, [&](dyb_ctx*, int r) -> int {
        int q = dyb_create(&t->m[r], t->dev[r], N, r * M, M);
        if (q || n_dev == 1) return q;
        if ((q = dyb_comm_init(t->m[r], r, n_dev, id))) return q;
        return dyb_comm_p2p_open_local(t->m[r], t->m.data(), 0);
    });
    if (!rc && n_dev > 1 && !(getenv("DYNEMOL_B200_P2P") && getenv("DYNEMOL_B200_P2P")[0] == '0'))
        rc = team_run(t, [&](dyb_ctx* c, int) -> int { return dyb_comm_p2p_open_local(c, t->m.data(), 1); });
    if (rc) { std::string e = t_err; dyb_team_destroy(t); return tfail(rc, e); }
    *out = t;
    return DYB_OK;
}

int dyb_team_form_hprime(dyb_team* t, const double* h_S, const double* h_h, double* h_H_out) {
    if (!t || !h_S || !h_h) return tfail(DYB_EINVAL, "NULL argument");
    int rc = ensure_form(t);
    if (rc) return rc;
    if (t->P == 1) {
        if ((rc = dyb_form_hprime_async(t->form, h_S, h_h, h_H_out)) || (rc = dyb_sync(t->form))) return tfail(rc, dyb_last_error());
        return DYB_OK;
    }
    const bool distributed = !(getenv("DYNEMOL_B200_TEAM_FORM") && !strcmp(getenv("DYNEMOL_B200_TEAM_FORM"), "single"));
    if (distributed) {
        // 1. every member uploads ITS block of columns of h over its own PCIe link while the first device factorises S
        int rc_factor = DYB_OK;
        std::string factor_err;
        rc = team_run(t, [&](dyb_ctx* c, int r) -> int {
            int q = dyb_upload_column_block(c, h_h);
            if (q) return q;
            if (r == 0) { rc_factor = dyb_factor_overlap(t->form, h_S); if (rc_factor) factor_err = dyb_last_error(); }
            return dyb_sync(c);
        });
        if (rc) return rc;
        if (rc_factor == DYB_OK) {
            void* dU = nullptr; int64_t ldu = 0;
            if ((rc = dyb_factor_device(t->form, &dU, &ldu))) return tfail(rc, dyb_last_error());
            // 2. S X = h[:, block] on every member with a copy of the factor (pulled over NVLink): 2 N^3 / P flops each;
            //    the solved block goes down to the host columns it belongs to, beside what follows
            if ((rc = team_run(t, [&](dyb_ctx* c, int) -> int { return dyb_solve_column_block(c, dU, ldu, h_H_out); }))) return rc;
            // 3. column blocks -> row blocks: every member pulls its M x M sub-blocks out of the peers' blocks
            std::vector<void*> X(t->P, nullptr);
            for (int r = 0; r < t->P; ++r) dyb_column_block_device(t->m[r], &X[r]);
            return team_run(t, [&](dyb_ctx* c, int) -> int { return dyb_take_rows_from_column_blocks(c, X.data(), t->P); });
        }
        if (rc_factor != DYB_ESINGULAR) return tfail(rc_factor, factor_err);
        // S is not numerically positive definite: the single-device route below has the LU fallback
    }
    if ((rc = dyb_form_hprime_async(t->form, h_S, h_h, h_H_out))) return tfail(rc, dyb_last_error());
    if ((rc = dyb_sync(t->form))) return tfail(rc, dyb_last_error());       // the solve is done; the download of H' may still run
    void* d = nullptr; int64_t ld = 0;
    dyb_hprime_device(t->form, &d, &ld);
    // every member pulls its row block out of the first device's H' (peer copy over NVLink)
    return team_run(t, [&](dyb_ctx* c, int) -> int { return dyb_upload_hprime_device(c, d, ld); });
}

int dyb_team_wait_outputs(dyb_team* t) {
    if (!t) return tfail(DYB_EINVAL, "team is NULL");
    int rc;
    for (dyb_ctx* c : t->m) if ((rc = dyb_wait_outputs(c))) return tfail(rc, dyb_last_error());
    if (t->form && (rc = dyb_wait_outputs(t->form))) return tfail(rc, dyb_last_error());
    return DYB_OK;
}

int dyb_team_upload_hprime(dyb_team* t, const double* h_H, int64_t lda) {
    if (!t || !h_H) return tfail(DYB_EINVAL, "NULL argument");
    return team_run(t, [&](dyb_ctx* c, int) -> int { return dyb_upload_hprime(c, h_H, lda); });
}

int dyb_team_set_packets(dyb_team* t, int n_part, const dyb_complex* bra, const dyb_complex* ket) {
    if (!t || !bra || !ket || n_part < 1 || n_part > 2) return tfail(DYB_EINVAL, "bad argument");
    int rc = team_run(t, [&](dyb_ctx* c, int) -> int { return dyb_set_packets(c, n_part, bra, ket); });
    if (!rc) t->n_part = n_part;
    return rc;
}

int dyb_team_get_packets(dyb_team* t, int n_part, dyb_complex* bra, dyb_complex* ket) {
    if (!t || !bra || !ket || n_part < 1 || n_part > 2) return tfail(DYB_EINVAL, "bad argument");
    const size_t n = (size_t)t->N * n_part;
    return team_run(t, [&](dyb_ctx* c, int r) -> int {          // collective: every rank gathers; rank 0 fills the caller's buffers
        if (r == 0) return dyb_get_packets(c, n_part, bra, ket);
        t->scratch[r].resize(2 * n);
        return dyb_get_packets(c, n_part, t->scratch[r].data(), t->scratch[r].data() + n);
    });
}

int dyb_team_set_spectral_bounds(dyb_team* t, double emin, double emax) {
    if (!t) return tfail(DYB_EINVAL, "team is NULL");
    for (dyb_ctx* c : t->m) { int rc = dyb_set_spectral_bounds(c, emin, emax); if (rc) return tfail(rc, dyb_last_error()); }
    return DYB_OK;
}

int dyb_team_estimate_spectral_bounds(dyb_team* t, int n_iter, double margin, double* emin, double* emax) {
    if (!t) return tfail(DYB_EINVAL, "team is NULL");
    std::vector<double> lo(t->P), hi(t->P);
    int rc = team_run(t, [&](dyb_ctx* c, int r) -> int { return dyb_estimate_spectral_bounds(c, n_iter, margin, &lo[r], &hi[r]); });
    if (rc) return rc;
    for (int r = 1; r < t->P; ++r)
        if (lo[r] != lo[0] || hi[r] != hi[0]) return tfail(DYB_ECUDA, "spectral bounds differ between the members (replicated arithmetic broken)");
    if (emin) *emin = lo[0];
    if (emax) *emax = hi[0];
    return DYB_OK;
}

int dyb_team_propagate(dyb_team* t, int mode, double t_init, double t_max, const double* tau, double* save_tau, dyb_trace* traces) {
    if (!t || !tau || !save_tau) return tfail(DYB_EINVAL, "NULL argument");
    std::vector<double> sv((size_t)2 * t->P, 0.0);
    int rc = team_run(t, [&](dyb_ctx* c, int r) -> int { return dyb_propagate(c, mode, t_init, t_max, tau, &sv[2 * r], r == 0 ? traces : nullptr); });
    if (rc) return rc;
    for (int r = 1; r < t->P; ++r)
        for (int p = 0; p < t->n_part; ++p)
            if (sv[2 * r + p] != sv[p]) return tfail(DYB_ECUDA, "tau schedules differ between the members (replicated decisions broken)");
    for (int p = 0; p < t->n_part; ++p) save_tau[p] = sv[p];
    return DYB_OK;
}

int dyb_team_ao_bra(dyb_team* t, int n_part, dyb_complex* h_AO_bra) {
    if (!t || !h_AO_bra || n_part < 1 || n_part > 2) return tfail(DYB_EINVAL, "bad argument");
    if (!t->form) return tfail(DYB_EINVAL, "dyb_team_ao_bra needs the factor of S: call dyb_team_form_hprime first");
    // S^-1 Psi_bra with the Cholesky factor kept on the first device: the gathered packets make a 64 N byte round trip
    const size_t n = (size_t)t->N * n_part;
    std::vector<dyb_complex> bra(n), ket(n);
    int rc = dyb_team_get_packets(t, n_part, bra.data(), ket.data());
    if (rc) return rc;
    if (t->P > 1 && (rc = dyb_set_packets(t->form, n_part, bra.data(), ket.data()))) return tfail(rc, dyb_last_error());
    if ((rc = dyb_ao_bra(t->form, n_part, h_AO_bra))) return tfail(rc, dyb_last_error());
    return DYB_OK;
}

int dyb_team_populations(dyb_team* t, int n_part, int n_frag, const int32_t* fragment, double tm, double* out) {
    if (!t || !fragment || !out || n_part < 1 || n_part > 2 || n_frag < 0) return tfail(DYB_EINVAL, "bad argument");
    std::vector<std::vector<double>> o(t->P, std::vector<double>((size_t)(n_frag + 2) * n_part));
    int rc = team_run(t, [&](dyb_ctx* c, int r) -> int { return dyb_populations(c, n_part, n_frag, fragment, tm, r == 0 ? out : o[r].data()); });
    return rc;
}

int dyb_team_quasiparticle_energies(dyb_team* t, int n_part, double* out_reim) {
    if (!t || !out_reim || n_part < 1 || n_part > 2) return tfail(DYB_EINVAL, "bad argument");
    std::vector<double> o((size_t)4 * t->P);
    int rc = team_run(t, [&](dyb_ctx* c, int r) -> int { return dyb_quasiparticle_energies(c, n_part, &o[4 * r]); });
    if (rc) return rc;
    for (int i = 0; i < 2 * n_part; ++i) out_reim[i] = o[i];
    return DYB_OK;
}

int dyb_team_run_terms(dyb_team* t, double tau, int n_terms, float* elapsed_ms_max) {
    if (!t) return tfail(DYB_EINVAL, "team is NULL");
    std::vector<float> ms(t->P, 0.f);
    int rc = team_run(t, [&](dyb_ctx* c, int r) -> int { return dyb_run_terms(c, tau, n_terms, &ms[r], nullptr); });
    if (rc) return rc;
    float mx = 0.f;
    for (float v : ms) mx = v > mx ? v : mx;
    if (elapsed_ms_max) *elapsed_ms_max = mx;
    return DYB_OK;
}

int64_t dyb_team_passes_last(dyb_team* t) {
    if (!t || t->m.empty()) return 0;
    int64_t info[16];
    if (dyb_get_info(t->m[0], info)) return 0;
    return info[11];
}

}  // extern "C"
