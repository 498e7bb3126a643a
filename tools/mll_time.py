import sys
sys.path.insert(0, ".")
import time, numpy as np, gpax_b200
ctx = gpax_b200.default_context()
rng = np.random.default_rng(0)
for N in (4096, 16384):
    X = rng.uniform(0, 1, (N, 3)); y = rng.standard_normal(N)
    th = np.array([0.3, 0.3, 0.3, 1.0, 0.1, 1.0])
    ctx.mll("RBF", X, y, th, 1e-6, want_grad=True)
    t0 = time.perf_counter()
    for _ in range(3): v, g, _, info = ctx.mll("RBF", X, y, th, 1e-6, want_grad=True)
    t1 = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    for _ in range(3): ctx.mll("RBF", X, y, th, 1e-6, want_grad=False)
    t2 = (time.perf_counter() - t0) / 3
    print(f"mll N={N}: value+grad {t1*1e3:.1f} ms, value only {t2*1e3:.1f} ms, info {info}, launches {ctx.last_timing()['launches']}", flush=True)
