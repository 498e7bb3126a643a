"""Factorisation / solve timing with the strip solve off, 256 and 512 wide (single stream, N = 16384 and 4096)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gpax_b200 import _ffi

ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
for N in (4096, 16384):
    P, d, S = 1024, 3, 2
    X = rng.uniform(0, 1, (N, d)); y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(N); Xn = rng.uniform(0, 1, (P, d))
    theta = np.tile(np.array([0.3, 0.3, 0.3, 1.0, 0.1, 1.0]), (S, 1))
    ref = None
    for strip in (0, 256, 512, 1024):
        ctx.set_option("streams", 1)
        ctx.set_option("trsm_strip", strip)
        for rep in range(2):
            o = ctx.posterior("RBF", X, y, Xn, theta, want=("mean", "var"), timing=True)
        t = o["timing"]
        if ref is None:
            ref = o
        err = np.abs(o["mean"] - ref["mean"]).max() / np.abs(ref["mean"]).max()
        print("N", N, "strip", strip, {k: round(v / S, 2) for k, v in t.items() if k in ("total_ms", "potrf_ms", "trsm_ms")}, "launches/draw",
              t["launches"] // S, "rel diff vs recursion", f"{err:.1e}", flush=True)
ctx.set_option("trsm_strip", 256)
