"""CPU suite: host-side logic of the product that needs no GPU -- the launch plan of the dual product (tile ranges,
segments, panel bookkeeping the fused epilogue relies on) and the synthetic workload generator."""
import numpy as np
import pytest

from dynemol_b200 import synthetic as syn


@pytest.fixture(scope="module")
def api():
    from dynemol_b200 import build
    build.build()
    from dynemol_b200 import api as a
    return a


@pytest.mark.parametrize("N,rows,sms", [(1, 1, 148), (5, 5, 148), (257, 257, 148), (2048, 2048, 148), (4100, 4100, 148),
                                        (10368, 10368, 148), (16384, 16384, 148), (65536, 8192, 148), (65536, 65536, 148),
                                        (30000, 30000, 132), (1000, 1000, 4)])
def test_plan_covers_every_tile_once_and_orders_segments_by_panel(api, N, rows, sms):
    p = api.plan(N, rows, sms)
    TC, PR = p["tile_cols"], p["panel_rows"]
    assert p["panels"] == -(-rows // PR) and p["tiles_per_panel"] == -(-N // TC)
    assert p["tiles"] == p["panels"] * p["tiles_per_panel"] and 1 <= p["grid"] <= min(sms, p["tiles"])
    if p["tiles"] >= 8 * sms:
        assert p["grid"] == sms, "large operators use every SM"
    assert p["padded_cols"] == p["tiles_per_panel"] * TC >= N
    G, T, TPP = p["grid"], p["tiles"], p["tiles_per_panel"]
    covered = 0; seg_panels = []
    for b in range(G):
        t0, t1 = T * b // G, T * (b + 1) // G
        assert t1 - t0 in (T // G, T // G + 1), "tile ranges are balanced to +-1"
        assert p["seg_base"][b] == len(seg_panels)
        covered += t1 - t0
        if t1 > t0:
            seg_panels += list(range(t0 // TPP, (t1 - 1) // TPP + 1))
    assert covered == T and len(seg_panels) == p["segments"]
    assert seg_panels == sorted(seg_panels), "segments must be globally ordered by panel (epilogue sums a contiguous range)"
    for q in range(p["panels"]):
        a, e = p["pseg_start"][q], p["pseg_start"][q + 1]
        assert e > a and all(x == q for x in seg_panels[a:e])
    assert p["pseg_start"][-1] == p["segments"]


def test_synthetic_workload_structure():
    w = syn.make_workload(128)
    assert np.array_equal(w.S, w.S.T) and np.allclose(np.diag(w.S), 1.0)
    ev = np.linalg.eigvalsh(w.S)
    assert ev[0] > 0 and ev[-1] / ev[0] < 1e3, "overlap must be SPD with cond(S) <= 1e3 (SURVEY.md 8d)"
    assert np.array_equal(w.h, w.h.T) and np.allclose(np.diag(w.h), w.IP)
    for p in range(2):                                      # packets: Psi_ket = C, Psi_bra = S C, <bra|ket> = 1
        assert np.allclose(w.Psi_bra[:, p], w.S @ w.Psi_ket[:, p])
        assert abs(np.vdot(w.Psi_bra[:, p], w.Psi_ket[:, p]) - 1.0) < 1e-12
    assert set(np.unique(w.fragment)) == {0, 1, 2, 3}
    pos, sp = syn.li2s_lattice(6, 6, 6)
    assert pos.shape == (2592, 3) and (sp == 1).sum() == 1728 and (sp == 0).sum() == 864     # examples/Li2S-crystal


def test_torch_generators_match_numpy():
    torch = pytest.importorskip("torch")
    N = 128
    w = syn.make_workload(N)
    S, h, _ = syn.make_S_h_torch(N, "cpu")
    assert np.abs(S.numpy() - w.S).max() < 1e-14 and np.abs(h.numpy() - w.h).max() < 1e-13
    shard = syn.make_h_shard_colmajor_torch(N, 32, 64, "cpu", dense_tail=False)
    assert np.abs(shard.numpy().T - w.h[32:96, :]).max() < 1e-13
    dense = syn.make_h_shard_colmajor_torch(N, 32, 64, "cpu")
    assert (dense == 0).sum() == 0 and np.abs(dense.numpy().T - w.h[32:96, :]).max() <= 1e-3


def test_resident_plan_blocking():
    """Host arithmetic of the shared-memory-resident series kernel (csrc/resident.cuh): the Gd x Gd blocking covers the
    operator, fits one CTA per SM and the opt-in shared memory, and keeps the column stride odd (bank-conflict-free
    from both sides).  Largest operator that fits on a B200: N = 1824 (152-row blocks on 12 x 12 CTAs)."""
    from dynemol_b200 import api
    for N in list(range(1, 70)) + [127, 128, 129, 383, 384, 385, 512, 892, 900, 1000, 1727, 1728, 1792, 1823, 1824]:
        p = api.resident_plan(N)
        assert p["fits"] == 1, N
        assert p["grid_side"] ** 2 <= 148 and p["grid_side"] * p["block"] >= N
        assert (p["grid_side"] - 1) * p["block"] < N, "no empty block row"
        assert p["smem_stride"] % 2 == 1 and p["smem_stride"] >= p["block"]
        assert p["block"] <= 152 and p["smem_bytes"] <= 227 * 1024 - 2048
        assert p["block"] >= min(N, 32) or p["grid_side"] == 12
        assert p["block"] * p["smem_stride"] * 8 < p["smem_bytes"]
    for N in (1825, 2048, 4096, 16384):
        assert api.resident_plan(N)["fits"] == 0
    assert api.resident_plan(1824)["grid_side"] == 12 and api.resident_plan(1824)["block"] == 152
    # a smaller GPU (fewer SMs) gets a smaller grid and therefore a smaller limit
    p = api.resident_plan(900, sm_count=64)
    assert p["grid_side"] == 8 and p["block"] == 113 and p["fits"] == 1
    assert api.resident_plan(1824, sm_count=64)["fits"] == 0


def test_mid_plan_blocking():
    """Host arithmetic of the streamed one-launch series kernel (csrc/mid.cuh): the Gr x Gc blocking covers the operator with
    no empty block row or column and at most one CTA per SM; block columns are whole tiles; every vector index has exactly ONE
    owner CTA (E consecutive indices each, spanning at most two block columns); at most 64 ket partials per entry; a
    consumer's R + Cnp entries fit 19 words per thread of warps 1..7; the ring, the union region (bra partials of the row groups, the
    staged ket rows, the partials and scalars an owner collects), the consumer copy of x and the owner state fit the opt-in shared memory."""
    from dynemol_b200 import api
    for N in [257, 300, 512, 520, 600, 640, 700, 1825, 2000, 2048, 2304, 2592, 3000, 3333, 4096, 4608, 5000, 5632, 6144]:
        p = api.mid_plan(N)
        assert p["fits"] == 1, (N, p)
        R, TC, Gr, Gc, Cnp = p["block_rows"], p["tile_cols"], p["grid_rows"], p["grid_cols"], p["block_cols"]
        E, n_own = p["owned"], p["owners"]
        assert R == 512 and TC == 8 and Cnp % TC == 0 and p["tiles_per_term"] == Cnp // TC
        assert Gr * Gc <= 148 and Gc <= 64
        assert Gr * R >= N > (Gr - 1) * R and Gc * Cnp >= N > (Gc - 1) * Cnp
        assert E * Gr * Gc >= N and n_own == -(-N // E) and n_own <= Gr * Gc and (n_own - 1) * E < N and E <= min(64, Cnp)
        assert p["collect_words"] == (Gr + Gc) * E * 4
        assert (Cnp + R) * 4 <= 19 * 224 and 2 * E * 4 <= 3 * 224 and p["collect_words"] <= 17 * 256 and 8 * n_own <= 5 * 256
        slots = -(-p["collect_words"] // 256)
        table = (-(-(Gr + Gc + 8 * E) * 4 // 128) * 128 + slots * 512) if p["table16"] else slots * 1024
        union = max(32768, -(-(p["collect_words"] + 8 * n_own) * 8 // 128) * 128)
        assert 2 * Cnp * 32 <= 32768 and R * 32 <= 32768
        assert 2 <= p["stages"] <= 5 and p["smem_bytes"] <= 227 * 1024 - 2048
        assert p["smem_bytes"] == p["stages"] * 32768 + 128 + union + 32 * (R + Cnp) + E * (4 * 2 * 32 + 32) + table
    assert api.mid_plan(4096)["grid_rows"] == 8 and api.mid_plan(4096)["grid_cols"] == 18 and api.mid_plan(4096)["block_cols"] == 232
    assert api.mid_plan(2048)["grid_cols"] == 37 and api.mid_plan(2048)["stages"] == 5 and api.mid_plan(2048)["owned"] == 14
    assert api.mid_plan(2048)["table16"] == 0 and api.mid_plan(6144)["table16"] == 1 and api.mid_plan(6144)["stages"] == 4
    for N in (7168, 8192, 16384):     # blocks wider than 512 columns: the consumer copy and the union region overflow;
        assert api.mid_plan(N)["fits"] == 0       # the two-launch path (>= 80 % of the HBM peak there) takes over


def test_mid_exchange_needs_no_barrier():
    """The barrier-free exchange of csrc/mid.cuh reuses its tagged buffers after two terms (scalars: four).  That is safe
    because of who waits for whom -- checked here on the blocking dyb_mid_plan returns, for the sizes the GPU tests run:
      * CTA (i, j) overwrites its partials of term t when it publishes term t+2, i.e. after it consumed x(t+2) on rows(i) and
        cols(j); the owners of those indices are exactly the readers of its partials of term t, and they published x(t+2)
        after reading them;
      * an owner overwrites x(t+1) when it publishes x(t+3), i.e. after it collected the partials of term t+2 of the block row
        and block column of its indices -- exactly the CTAs that read its x(t+1);
      * the product of a term depends on the products of EVERY CTA two terms earlier (two hops of "the producers of the
        entries I consume"), which is what makes a four-deep ring enough for the scalars every CTA reads."""
    from dynemol_b200 import api
    for N in [257, 300, 700, 2048, 2304, 3000, 4096, 5000, 6144]:
        p = api.mid_plan(N)
        R, Gr, Gc, Cnp, E = p["block_rows"], p["grid_rows"], p["grid_cols"], p["block_cols"], p["owned"]
        G = Gr * Gc
        rows = lambda i: range(i * R, min((i + 1) * R, N))
        cols = lambda j: range(j * Cnp, min((j + 1) * Cnp, N))
        producers = {}
        for b in range(G):
            i, j = divmod(b, Gc)
            readers = {g // E for g in rows(i)} | {g // E for g in cols(j)}       # owners that collect b's ket / bra partials
            awaited = {g // E for g in rows(i)} | {g // E for g in cols(j)}       # owners of the entries b consumes next
            assert readers <= awaited
            d = set()                                                             # CTAs whose partials feed those entries
            for bc in {g // Cnp for g in rows(i)}:
                d |= {r * Gc + bc for r in range(Gr)}
            for br in {g // R for g in cols(j)}:
                d |= {br * Gc + c for c in range(Gc)}
            producers[b] = d
        for o in range(p["owners"]):
            idx = range(o * E, min((o + 1) * E, N))
            x_readers = {(g // R) * Gc + c for g in idx for c in range(Gc)} | {r * Gc + g // Cnp for g in idx for r in range(Gr)}
            collected = {b for b in range(G) if o in ({g // E for g in rows(b // Gc)} | {g // E for g in cols(b % Gc)})}
            assert x_readers <= collected, (N, o)
        for b in range(G):
            assert set().union(*[producers[d] for d in producers[b]]) == set(range(G)), (N, b)


@pytest.mark.parametrize("name", ["prop_N64_dt5e-6", "prop_N128_dt2e-5", "cheb_N64_dt5e-4", "cheb_N128_dt5e-5"])
def test_steady_schedule_matches_oracle_traces(golden_dir, name):
    """dyb_steady_schedule (the sub-step schedule the library predicts when it chains the steady loop of Taylor.f:81-126
    into one launch) against the tau the oracle recorded for every steady sub-step of the golden runs: after the first
    successful Convergence() (Taylor.f:65-78), as long as no norm test fails, the sequences must be identical."""
    import os
    from dynemol_b200 import api
    H_BAR = 6.58264e-4
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    t_init, t_max = float(g["t_init"]), float(g["t_max"])
    checked = 0
    for tag in ("el", "hl"):
        ev = g[f"{tag}_events"]; etau = g[f"{tag}_event_tau"]
        i0 = next(i for i in range(len(ev)) if ev[i, 0] == 1 and ev[i, 2] == 1)        # first Convergence that succeeded
        tau = float(etau[i0])
        t = t_init + tau * H_BAR                                                      # Taylor.f:73
        if t_max - t < tau * H_BAR:                                                   # Taylor.f:75-78
            tau = (t_max - t) / H_BAR
        steady = []
        for i in range(i0 + 1, len(ev)):
            if ev[i, 0] != 2 or ev[i, 2] != 1:
                break
            steady.append(float(etau[i]))
        complete = (i0 + 1 + len(steady) == len(ev))                                  # no failure until the end of the step
        pred = api.steady_schedule(t, t_max, tau)
        assert len(pred) >= len(steady)
        assert np.array_equal(pred[: len(steady)], np.array(steady)), (tag, pred[:4], steady[:4])
        if complete and len(ev) < 256:                                                # (the trace keeps 256 events)
            assert len(pred) == len(steady)
        checked += len(steady)
    assert checked > 0


def test_steady_schedule_edges():
    from dynemol_b200 import api
    H_BAR = 6.58264e-4
    assert len(api.steady_schedule(1.0, 1.0, 0.1)) == 0                               # t == t_max: nothing left (Taylor.f:81)
    s = api.steady_schedule(0.0, 10 * 0.01 * H_BAR, 0.01)
    assert 10 <= len(s) <= 11 and np.all(s[:9] == 0.01) and s.sum() * H_BAR == pytest.approx(10 * 0.01 * H_BAR, rel=1e-12)
    assert len(api.steady_schedule(0.0, 1.0, 1e-3, max_sub=7)) == 7                    # capped


def test_series_coefficients_match_oracle(oracle_mod):
    """The coefficient tables and k_max rules of the library's host logic (dyb_series_coefficients) against the oracle
    (Taylor.f:224-239, :165-171; Chebyshev_gpu.cpp:636-643, :565-574 with the spectral rescaling) and the independent
    numpy transcription."""
    from dynemol_b200 import api
    from oracle import taylor_numpy as tn
    for tau in (1e-6, 7.6e-5, 3.0e-3, 0.03, 0.76):
        C, km = api.series_coefficients(api.MODE_TAYLOR, tau)
        ref = oracle_mod.coefficient(tau)
        assert np.array_equal(C, ref) or np.abs(C - ref).max() <= 1e-16 * np.abs(ref).max()
        assert np.abs(C - tn.coefficient(tau)).max() <= 1e-16
        small = np.nonzero(np.abs(ref[1:]) < 1.0e-16)[0]
        assert km == (int(small[0]) + 2 if small.size else 25)                         # Taylor.f:165-171
    for tau, ebar, de in ((0.076, 480.0, 560.0), (0.76, 500.0, 520.0), (7.6e-3, 10.0, 30.0)):
        C, km = api.series_coefficients(api.MODE_CHEBYSHEV, tau, ebar, de)
        ref = oracle_mod.cheb_scaled_coefficient(tau, ebar, de)
        assert np.abs(C - ref).max() <= 4e-16 and np.abs(C - tn.cheb_coefficient(tau, ebar, de)).max() <= 4e-16
        R = de * tau
        want = 25
        for k in range(6, 25):                                                         # Chebyshev_gpu.cpp:565-574
            if abs(ref[k] * oracle_mod.naked_bessel(k, R)) < 1.0e-20:
                want = k
                break
        assert km == want
    with pytest.raises(Exception):
        api.series_coefficients(api.MODE_CHEBYSHEV, 0.1, 0.0, 0.0)


def test_gpu_pin_size_recovery():
    """gpu_pin_ receives n*n*8 as a Fortran DEFAULT integer (ElHl_Chebyshev_GPU.f:109-111): it wraps at N = 16384 and
    loses multiples of 2^32 beyond N = 23170.  The library recovers the true byte count (N x N doubles)."""
    from dynemol_b200 import api
    for N in (64, 900, 10368, 16384, 23170, 23171, 30720, 32768, 46340, 60000):
        true = N * N * 8
        wrapped = ((true + 2 ** 31) % 2 ** 32) - 2 ** 31          # what a 32-bit signed integer holds
        assert api.unwrap_pin_bytes(wrapped) == true, N
    # inherent ambiguity of the wrapped value: N = 65536 and N = 32768 both arrive as 0; the smaller size is pinned (the
    # rest of such a buffer stays pageable: slower copies, still correct)
    assert api.unwrap_pin_bytes(0) == 32768 * 32768 * 8
