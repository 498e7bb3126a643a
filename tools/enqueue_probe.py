"""How far ahead of the GPU does the host get while queueing a multi-draw posterior?  (host_enqueue_ms vs total_ms)"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
sys.path.insert(0, ".")
from gpax_b200 import _ffi

ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
N, P, d, S = 16384, 1024, 3, 16
X = rng.uniform(0, 1, (N, d)); y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(N); Xn = rng.uniform(0, 1, (P, d))
theta = np.tile(np.array([0.3, 0.3, 0.3, 1.0, 0.1, 1.0]), (S, 1))
for streams, thr in ((1, 1), (2, 1), (4, 0), (4, 1), (8, 1)):
    ctx.set_option("streams", streams)
    ctx.set_option("enqueue_threads", thr)
    for rep in range(2):
        t0 = time.perf_counter()
        o = ctx.posterior("RBF", X, y, Xn, theta, want=("mean", "var"), timing=False)
        wall = (time.perf_counter() - t0) * 1e3
    t = ctx.last_timing()
    print("streams", streams, "threads", thr, "total_ms", round(t["total_ms"], 1), "per draw", round(t["total_ms"] / S, 2), "host_enqueue_ms",
          round(t["host_enqueue_ms"], 1), "launches", t["launches"], flush=True)
