"""potrf_once.py N [reps] -- one device-resident Gram + factorisation (for ncu launch lists)."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from gpax_b200 import _ffi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
d = 3
X = ctx.to_device(rng.uniform(0, 1, (N, d)))
ell = np.full(d, 0.3)
K = ctx.alloc((N, N))
for _ in range(reps):
    ctx._check(ctx.lib.b2gp_gram(ctx.h, 0, X.ptr, N, X.ptr, N, d, _ffi._ptr(ell), 1.0, 1.0, 0.1 + 1e-6, 1, K.ptr, N,
                                 _ffi.FLAG_DEVICE_PTRS | _ffi.FLAG_LOWER_ONLY))
    info = C.c_int(0)
    ctx._check(ctx.lib.b2gp_potrf(ctx.h, N, K.ptr, N, C.byref(info), _ffi.FLAG_DEVICE_PTRS))
    t = ctx.last_timing()
    print(f"N={N} potrf {t['total_ms']:.3f} ms  {N**3/3/t['total_ms']/1e9:.2f} TF/s  launches {t['launches']} info {info.value}")
