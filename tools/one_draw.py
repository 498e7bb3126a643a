"""Two single-draw posteriors at the headline size on one stream (for `ncu` launch lists: the second draw is warm)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from gpax_b200 import _ffi

ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
N, P, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 16384), 1024, 3
X = rng.uniform(0, 1, (N, d)); y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(N); Xn = rng.uniform(0, 1, (P, d))
ctx.set_option("streams", 1)
for rep in range(2):
    theta = np.array([[0.3, 0.3, 0.3, 1.0 + 0.01 * rep, 0.1, 1.0]])
    o = ctx.posterior("RBF", X, y, Xn, theta, want=("mean", "var"))
    t = ctx.last_timing()
    print(rep, "total_ms", round(t["total_ms"], 2), "launches", t["launches"], flush=True)
