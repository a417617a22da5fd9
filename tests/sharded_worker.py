"""Worker of the multi-GPU parity test (launched by torchrun, one rank per GPU): row-sharded
propagation through the C ABI compared with the CPU oracle on rank 0."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import numpy as np
import torch                      # before dynemol_b200.api: the library binds the NCCL copy torch loaded
import torch.distributed as dist


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DYB_TEST_WATCHDOG", "240")), exit=True)   # never hang a GPU box
    from dynemol_b200 import api, sharded, synthetic as syn
    rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
    verbose = os.environ.get("DYB_TEST_VERBOSE") == "1"

    def mark(msg):
        if verbose:
            print(f"[rank {rank}] {msg}", flush=True)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, dt = int(os.environ.get("DYB_TEST_N", "1024")), float(os.environ.get("DYB_TEST_DT", "1e-6"))
    w = syn.make_workload(N)
    Hp = np.asfortranarray(np.linalg.solve(w.S, w.h))
    mark('workload built')
    P, row0, m = sharded.init_sharded(N, dist, local)
    mark('comm ready')
    P.upload_hprime(Hp)                              # keeps rows row0..row0+m only
    P.set_packets(w.Psi_bra, w.Psi_ket)
    mark('packets set')
    tau0 = dt / api.H_BAR
    save_tau, traces = P.propagate(0.0, dt, tau0)
    mark('propagated')
    bra, ket = P.get_packets()
    mark('packets read')
    out = {"rank": rank, "ok": True, "p2p": int(P.info()["p2p"])}
    if rank == 0:
        import oracle
        worst = 0.0
        for p in range(2):
            b, k, _, st, tr = oracle.propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt, tau0)
            eb = np.abs(bra[:, p] - b).max() / np.abs(b).max(); ek = np.abs(ket[:, p] - k).max() / np.abs(k).max()
            worst = max(worst, eb, ek)
            same_trace = [(e[0], e[1], e[2]) for e in traces[p].events()] == [(e[0], e[1], e[2]) for e in tr.events()]
            out["ok"] = bool(out["ok"] and same_trace and save_tau[p] == st and eb < 1e-10 and ek < 1e-10)
        out["worst_rel_err"] = float(worst)
        out["pairs"] = [int(t.n_matvec_pairs) for t in traces]
    dist.barrier()

    # Chebyshev mode through the same exchange (three-term recurrence reads x and x_prev on the owned slice)
    e = np.linalg.eigvals(Hp).real
    lo, hi = e.min() - 0.05 * (e.max() - e.min()), e.max() + 0.05 * (e.max() - e.min())
    dt2 = 20 * dt; tau2 = dt2 / api.H_BAR
    P.set_packets(w.Psi_bra, w.Psi_ket)
    P.set_spectral_bounds(lo, hi)
    save2, traces2 = P.propagate(0.0, dt2, tau2, mode=api.MODE_CHEBYSHEV)
    bra2, ket2 = P.get_packets()
    mark('chebyshev propagated')
    if rank == 0:
        worst2 = 0.0
        for p in range(2):
            b, k, _, st, tr = oracle.cheb_scaled_propagation(Hp, w.Psi_bra[:, p], w.Psi_ket[:, p], 0.0, dt2, tau2, 0.5 * (hi + lo), 0.5 * (hi - lo))
            eb = np.abs(bra2[:, p] - b).max() / np.abs(b).max(); ek = np.abs(ket2[:, p] - k).max() / np.abs(k).max()
            worst2 = max(worst2, eb, ek)
            same = [(x[0], x[1], x[2]) for x in traces2[p].events()] == [(x[0], x[1], x[2]) for x in tr.events()]
            out["ok"] = bool(out["ok"] and same and eb < 1e-10 and ek < 1e-10)
        out["worst_rel_err_cheb"] = float(worst2)
        out["pairs_cheb"] = [int(t.n_matvec_pairs) for t in traces2]
        print("SHARDED_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    P.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
