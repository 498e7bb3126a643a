"""bench_configs.py -- the single-GPU configurations of BASELINE.json other than the headline, through the public API
(host arrays in, host arrays out), each checked by a size-independent property:
   c1  ExactGP RBF 1D, N=512 d=1, single draw (the CPU-parity case; compared with the oracle outright)
   c2  ExactGP Matern52 2D, N=8192 d=2, 200-draw batched predict (mean + diagonal variance per draw)
   c3  viGP image reconstruction, N_train=16384 d=2, the whole 181x181 grid predicted in ONE call and in 1000-point chunks
prints one JSON line per config."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import gpax_b200  # noqa: E402
import oracle  # noqa: E402  (checker only)


def c1():
    rng = np.random.default_rng(0)
    N, P = 512, 1024
    X = rng.uniform(0, 1, (N, 1))
    y = np.sin(6 * X[:, 0]) + 0.1 * rng.standard_normal(N)
    Xn = np.linspace(0, 1, P)[:, None]
    params = {"k_length": np.array([0.2]), "k_scale": 1.0, "noise": 0.1}
    m = gpax_b200.ExactGP(1, "RBF")
    m.X_train, m.y_train = X, y
    m.get_mvn_posterior(Xn, params)
    t0 = time.perf_counter()
    for _ in range(5):
        m.ctx.set_option("drop_factor_cache", 1)
        mean, cov = m.get_mvn_posterior(Xn, params)
    dt = (time.perf_counter() - t0) / 5
    t1 = time.perf_counter()
    rm, rc = oracle.exact_posterior(X, y, Xn, params, "RBF")
    cpu = time.perf_counter() - t1
    err = max(np.abs(mean - rm).max() / np.abs(rm).max(), np.abs(cov - rc).max() / np.abs(rc).max())
    return {"config": "c1 ExactGP RBF N=512 d=1 P=1024 full covariance", "ms": dt * 1e3, "cpu_oracle_ms": cpu * 1e3,
            "max_scaled_err_vs_oracle": err, "cond_K": float(np.linalg.cond(oracle.rbf_kernel(X, X, params, 0.1)))}


def c2():
    rng = np.random.default_rng(1)
    N, P, d, S = 8192, 1024, 2, 200
    X = rng.uniform(0, 1, (N, d))
    y = np.sin(4 * X[:, 0]) * np.cos(3 * X[:, 1]) + 0.1 * rng.standard_normal(N)
    Xn = rng.uniform(0, 1, (P, d))
    r2 = np.random.default_rng(2)
    samples = {"k_length": np.exp(r2.normal(np.log(0.3), 0.1, (S, d))), "k_scale": np.exp(r2.normal(0, 0.1, S)),
               "noise": np.exp(r2.normal(np.log(0.1), 0.1, S))}
    theta = np.concatenate([samples["k_length"], samples["k_scale"][:, None], samples["noise"][:, None], np.ones((S, 1))], 1)
    ctx = gpax_b200.default_context()
    ctx.set_option("streams", 8)
    ctx.posterior("Matern", X, y, Xn, theta[:4], want=("mean", "var"))
    t0 = time.perf_counter()
    out = ctx.posterior("Matern", X, y, Xn, theta, want=("mean", "var"), timing=True)
    dt = time.perf_counter() - t0
    # property: draw s computed alone == draw s inside the batch (bitwise), and var > 0, info == 0
    one = ctx.posterior("Matern", X, y, Xn, theta[17:18], want=("mean", "var"))
    same = bool(np.array_equal(one["mean"][0], out["mean"][17]) and np.array_equal(one["var"][0], out["var"][17]))
    fl = S * (N ** 3 / 3 + N * N * (P + 1))
    return {"config": "c2 ExactGP Matern52 N=8192 d=2, 200-draw batched predict, P=1024 (mean + diag var)", "s": dt,
            "posteriors_per_s": S / dt, "tflops": fl / dt / 1e12, "info_all_zero": bool((out["info"] == 0).all()),
            "var_positive": bool((out["var"] > 0).all()), "batched_equals_single_bitwise": same,
            "ymean_absmax": float(np.abs(out["mean"].mean(0)).max())}


def c3():
    rng = np.random.default_rng(3)
    side = 181
    grid = np.stack(np.meshgrid(np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 2).astype(float)
    idx = rng.choice(side * side, 16384, replace=False)
    X = grid[idx]
    f = lambda g: 0.5 + 0.3 * np.sin(g[:, 0] / 14.0) * np.cos(g[:, 1] / 11.0) + 0.1 * np.sin((g[:, 0] + g[:, 1]) / 23.0)  # noqa: E731
    y = f(X) + 0.02 * rng.standard_normal(len(X))
    params = {"k_length": np.array([4.2, 3.2]), "k_scale": 0.05, "noise": 0.002}
    m = gpax_b200.viGP(2, "Matern")
    m.X_train, m.y_train = X, y
    m.predict(0, grid[:100], params)
    m.ctx.set_option("drop_factor_cache", 1)
    t0 = time.perf_counter()
    mean, var = m.predict(0, grid, params, noiseless=True)          # 32761 test points in one call
    t_one = time.perf_counter() - t0
    m.ctx.set_option("drop_factor_cache", 1)
    t0 = time.perf_counter()
    mb, vb = m.predict_in_batches(0, grid, 1000, params, noiseless=True)   # the reference notebook's chunking
    t_chunks = time.perf_counter() - t0
    rmse = float(np.sqrt(np.mean((mean - f(grid)) ** 2)))
    return {"config": "c3 viGP Matern52 N_train=16384 d=2 (181x181 image, 50% observed), full-grid reconstruction P=32761",
            "one_call_s": t_one, "chunks_of_1000_s": t_chunks, "chunked_equals_one_call": bool(np.allclose(mb, mean, rtol=1e-10, atol=1e-12)
                                                                                             and np.allclose(vb, var, rtol=1e-9, atol=1e-12)),
            "reconstruction_rmse": rmse, "var_min": float(var.min()), "var_max": float(var.max())}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3"]
    for w in which:
        print(json.dumps({"c1": c1, "c2": c2, "c3": c3}[w]()))
