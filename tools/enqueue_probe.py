"""How far ahead of the GPU does the host get while queueing a multi-draw posterior?  (host_enqueue_ms vs total_ms)"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gpax_b200 import _ffi

ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
N, P, d, S = 16384, 1024, 3, 16
X = rng.uniform(0, 1, (N, d)); y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(N); Xn = rng.uniform(0, 1, (P, d))
theta = np.tile(np.array([0.3, 0.3, 0.3, 1.0, 0.1, 1.0]), (S, 1))
for streams, thr, bg in ((4, 0, 0), (4, 1, 0), (4, 1, 140), (4, 1, 132), (4, 1, 120), (8, 1, 0), (8, 1, 132), (8, 0, 132), (6, 1, 132)):
    ctx.set_option("streams", streams)
    ctx.set_option("enqueue_threads", thr)
    ctx.set_option("big_grid", bg)
    for rep in range(2):
        t0 = time.perf_counter()
        o = ctx.posterior("RBF", X, y, Xn, theta, want=("mean", "var"), timing=False)
        wall = (time.perf_counter() - t0) * 1e3
    t = ctx.last_timing()
    print("streams", streams, "threads", thr, "big_grid", bg, "total_ms", round(t["total_ms"], 1), "host_enqueue_ms", round(t["host_enqueue_ms"], 1), "wall", round(wall, 1),
          "launches", t["launches"], flush=True)
