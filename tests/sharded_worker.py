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
        print("SHARDED_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    P.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
