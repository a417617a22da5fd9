"""GPU-box script (run under compute-sanitizer): a few short series of every series kernel at ragged sizes.

    compute-sanitizer --tool memcheck python tools/gpu_mid_sanitize.py 700 2304   # each N through mid, term and (N <= 1824) resident
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dynemol_b200 import api
for N in [int(a) for a in sys.argv[1:]] or [700, 2304]:
    g = torch.Generator(device="cuda").manual_seed(N)
    H = torch.randn((N, N), device="cuda", dtype=torch.float64, generator=g) / np.sqrt(N)
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((N, 2)) + 1j * rng.standard_normal((N, 2))) / np.sqrt(N)
    for kind in ["mid", "term"] + (["resident"] if N <= 1824 else []):
        P = api.Propagator(N)
        P.set_series_kernel(kind)
        P.upload_hprime_device(H.data_ptr(), N)
        P.set_packets(x, x.conj())
        ms, _ = P.run_terms(1e-4, 48)
        b, k = P.get_packets()
        print("N", N, kind, "ok", np.isfinite(b).all() and np.isfinite(k).all(), "ms", round(ms, 2), flush=True)
        P.close()
