"""GPU-box script: us per el+hole term of the mid-size series kernel (csrc/mid.cuh) against the two-launch path.

    python tools/gpu_mid.py [--sizes 2048,3000,4096,6144] [--l2mb 0,64,96,110] [--series 24] [--reps 100]

Every (N, L2 budget) pair gets its own context (DYNEMOL_B200_MID_L2MB is read when the context is created).  H' is a dense
random operator scaled to unit spectral radius (only throughput is measured); a term is one pass over H' for the four
right-hand sides plus the fused update, exactly what bench.py counts."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2048,3000,4096,6144")
    ap.add_argument("--l2mb", default="96")
    ap.add_argument("--series", type=int, default=24)
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from dynemol_b200 import api
    rows = []
    for N in [int(x) for x in args.sizes.split(",")]:
        g = torch.Generator(device="cuda").manual_seed(N)
        H = torch.randn((N, N), device="cuda", dtype=torch.float64, generator=g) / np.sqrt(N)
        rng = np.random.default_rng(1)
        x = (rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))) / np.sqrt(N)
        row = {"N": N}
        for kind, l2 in [("term", None)] + [("mid", v) for v in args.l2mb.split(",")]:
            if l2 is not None:
                os.environ["DYNEMOL_B200_MID_L2MB"] = l2
            P = api.Propagator(N)
            try:
                P.set_series_kernel(kind)
            except Exception as e:
                row["mid_error"] = str(e)[:120]
                P.close()
                continue
            P.upload_hprime_device(H.data_ptr(), N)
            P.set_packets(x, x.conj())
            for _ in range(3):
                P.run_terms(1e-4, args.series)
            ms, _ = P.run_terms(1e-4, args.series * args.reps)
            us = ms * 1e3 / (args.series * args.reps)
            key = "per_term_us" if kind == "term" else "mid_us_l2mb_%s" % l2
            row[key] = round(us, 2)
            row.setdefault("series_kernel", {})[key] = P.info()["series_kernel"]
            P.close()
        row["hbm_floor_us"] = round(8.0 * N * N / 6.5e12 * 1e6, 2)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
