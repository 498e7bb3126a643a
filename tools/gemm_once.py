"""gemm_once.py [n] [lower] [ozaki] -- a few device-resident launches of the rank-k update C -= A A^T (for ncu --set full):
ozaki = 0 the fp64 DMMA kernel (gemm_tma_kernel), 6 / 7 the int8 tcgen05 path (oz_slice_kernel + oz_mma_kernel)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from gpax_b200 import _ffi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lower = int(sys.argv[2]) if len(sys.argv) > 2 else 1
oz = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = _ffi.Context(0)
ctx.set_option("ozaki", oz)
rng = np.random.default_rng(0)
A = ctx.to_device(rng.standard_normal((n, n)))
Cm = ctx.to_device(np.zeros((n, n)))
for _ in range(3):
    ctx._check(ctx.lib.b2gp_gemm_nt(ctx.h, n, n, n, -1.0, A.ptr, n, A.ptr, n, 1.0, Cm.ptr, n, lower, _ffi.FLAG_DEVICE_PTRS))
    t = ctx.last_timing()
    print(f"n={n} lower={lower}: {t['epilogue_ms']:.3f} ms  {(1 if lower else 2) * n**3 / t['epilogue_ms'] / 1e9:.2f} TF/s")
