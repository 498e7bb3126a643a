"""gram_once.py [N] [kind] -- device-resident Gram builds (for ncu / timing)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from gpax_b200 import _ffi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
d = 3
X = ctx.to_device(rng.uniform(0, 1, (N, d)))
ell = np.full(d, 0.3)
K = ctx.alloc((N, N))
for lower in (0, 1):
    for rep in range(3):
        ctx._check(ctx.lib.b2gp_gram(ctx.h, kind, X.ptr, N, X.ptr, N, d, _ffi._ptr(ell), 1.0, 0.7, 0.1 + 1e-6, 1, K.ptr, N,
                                     _ffi.FLAG_DEVICE_PTRS | (_ffi.FLAG_LOWER_ONLY if lower else 0)))
    ms = ctx.last_timing()["total_ms"]
    nbytes = 8.0 * N * N * (0.5 if lower else 1.0)
    print(f"kind={kind} N={N} lower={lower}: {ms:.3f} ms  {nbytes/ms/1e6:.0f} GB/s")
