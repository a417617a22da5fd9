"""GPU-box script: FP64 rates of the building blocks of H' = S^-1 h at N = 16384 (what bounds dyb_form_hprime)."""
import time, torch
N = 16384
torch.manual_seed(0)
A = torch.randn(N, N, device="cuda", dtype=torch.float64)
S = A @ A.T / N + torch.eye(N, device="cuda", dtype=torch.float64) * 2.0
h = torch.randn(N, N, device="cuda", dtype=torch.float64)
del A
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    t = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    return min(t), r
t, U = timeit(lambda: torch.linalg.cholesky(S, upper=True)); print("potrf        %.3f s  %.1f TFLOP/s" % (t, N**3 / 3 / t / 1e12))
t, X = timeit(lambda: torch.cholesky_solve(h, U, upper=True)); print("potrs (N rhs) %.3f s  %.1f TFLOP/s" % (t, 2 * N**3 / t / 1e12))
t, Y = timeit(lambda: torch.linalg.solve_triangular(U, h, upper=True)); print("trsm (one)   %.3f s  %.1f TFLOP/s" % (t, N**3 / t / 1e12))
t, G = timeit(lambda: h @ S); print("dgemm        %.3f s  %.1f TFLOP/s" % (t, 2 * N**3 / t / 1e12))
t, Ui = timeit(lambda: torch.linalg.solve_triangular(U, torch.eye(N, device="cuda", dtype=torch.float64), upper=True)); print("trtri via trsm(I) %.3f s" % t)
t, Z = timeit(lambda: Ui @ (Ui.T @ h)); print("two dgemm with U^-1  %.3f s" % t)
print("max |Z - X| / max|X| = %.2e" % ((Z - X).abs().max() / X.abs().max()).item())
