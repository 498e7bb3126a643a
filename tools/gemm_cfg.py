import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
from gpax_b200 import _ffi
ctx = _ffi.Context(0)
fn = ctx.lib.b2gp_debug_gemm_cfg
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int] + [C.c_int64] * 3 + [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double)]
rng = np.random.default_rng(0)
n = 8192
A = ctx.to_device(rng.standard_normal((n, n)))
Cm = ctx.to_device(np.zeros((n, n)))
for (m_, n_, k_, lower) in ((8192, 8192, 8192, 1), (8192, 8192, 8192, 0), (8192, 8192, 512, 1), (8192, 8192, 128, 1)):
    for cfg in range(6):
        best = 1e9
        for rep in range(3):
            ms = C.c_double()
            rc = fn(ctx.h, cfg, m_, n_, k_, A.ptr, n, A.ptr, n, Cm.ptr, n, lower, C.byref(ms))
            assert rc == 0, ctx.lib.b2gp_last_error(ctx.h)
            best = min(best, ms.value)
        fl = (1 if lower else 2) * m_ * n_ * k_
        print(f"m={m_} k={k_} lower={lower} cfg={cfg}: {best:8.3f} ms {fl/best/1e9:6.2f} TF/s")
