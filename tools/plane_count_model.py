"""plane_count_model.py -- how many digit planes does the posterior need?  CPU model built on the bit-level
restatement of the int8 GEMM (oracle/ozaki_oracle.py): a blocked right-looking Cholesky whose trailing updates all go
through the digit-plane GEMM with S planes (pessimistic: on the GPU only updates with k >= 512 do), then the posterior
mean / variance by triangular solves, compared with the all-fp64 result and with the 1e-9 parity bar.
 usage: python tools/plane_count_model.py [N] [block]"""
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, ".")
import oracle                                      # noqa: E402  (checker only)
from oracle import ozaki_oracle as oz             # noqa: E402


def chol_blocked(K, nb, S):
    A = K.copy()
    n = A.shape[0]
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        A[j0:j1, j0:j1] = sla.cholesky(A[j0:j1, j0:j1], lower=True)
        if j1 < n:
            A[j1:, j0:j1] = sla.solve_triangular(A[j0:j1, j0:j1], A[j1:, j0:j1].T, lower=True).T
            P = A[j1:, j0:j1]
            if S == 0:
                A[j1:, j1:] -= P @ P.T
            else:
                A[j1:, j1:] = oz.gemm_nt(P, P, A[j1:, j1:], alpha=-1.0, S=S)
    return np.tril(A)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (N, 3))
    y = np.sin(3 * X[:, 0]) + X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(N)
    Xn = rng.uniform(0, 1, (64, 3))
    print(f"N={N} block={nb}; error = max |x - x_fp64| / max |x_fp64| of the posterior mean and variance")
    print(f"{'noise':>8} {'cond(K)':>9} | " + " | ".join(f"S={S}: mean     var   " for S in (8, 7, 6)))
    for noise in (0.1, 1e-2, 1e-3, 1e-4, 1e-5):
        params = {"k_length": np.full(3, 0.3), "k_scale": 1.0, "noise": noise}
        K = oracle.rbf_kernel(X, X, params, noise)
        kpx = oracle.rbf_kernel(Xn, X, params, jitter=0.0)
        cond = np.linalg.cond(K)
        res = {}
        for S in (0, 8, 7, 6):
            L = chol_blocked(K, nb, S)
            V = sla.solve_triangular(L, kpx.T, lower=True)
            w = sla.solve_triangular(L, y, lower=True)
            res[S] = (V.T @ w, 1.0 + noise + 1e-6 - (V * V).sum(0))
        row = []
        for S in (8, 7, 6):
            em = np.abs(res[S][0] - res[0][0]).max() / np.abs(res[0][0]).max()
            ev = np.abs(res[S][1] - res[0][1]).max() / np.abs(res[0][1]).max()
            row.append(f"{em:8.1e} {ev:8.1e}")
        print(f"{noise:8.0e} {cond:9.1e} | " + " | ".join(row))


if __name__ == "__main__":
    main()
