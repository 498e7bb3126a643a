"""i8_peak.py -- measured int8 tcgen05 ceiling (TOP/s) of this GPU: bare UTCIMMA loop, no TMA, no epilogue."""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
from gpax_b200 import _ffi  # noqa: E402

ctx = _ffi.Context(0)
fn = ctx.lib.b2gp_debug_i8_peak
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
out = {}
for iters in (2048, 16384, 131072):
    tops, ms = C.c_double(), C.c_double()
    rc = fn(ctx.h, iters, 5, C.byref(tops), C.byref(ms))
    assert rc == 0, ctx.lib.b2gp_last_error(ctx.h)
    out[str(iters)] = {"int8_tops": tops.value, "ms": ms.value}
print(json.dumps({"kernel": "oz_i8_peak_kernel (tcgen05.mma kind::i8 M128 N256 K32, operands resident in shared memory)", "by_iters": out}))
