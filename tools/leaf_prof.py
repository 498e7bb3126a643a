import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
from gpax_b200 import _ffi
ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
n = 128
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
A = (Q * np.geomspace(1, 100, n)) @ Q.T
A = (A + A.T) / 2
dA = ctx.to_device(A)
dL = ctx.alloc((128, 128))
fn = ctx.lib.b2gp_debug_leaf
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
prof = np.zeros(64, dtype=np.int64)
names = ["start", "loaded", "a0_loadD", "a0_factor", "a0_writeL", "a0_inverse", "a0", "bc0", "a1", "bc1", "a2", "bc2", "a3", "bc3", "Lwritten", "inv_assembled", "end"]
for rep in range(3):
    dA.upload(A)
    ctx._check(fn(ctx.h, n, dA.ptr, n, dL.ptr, prof.ctypes.data))
    cyc, ns = prof[0::2], prof[1::2]
    k = len(names)
    print("rep", rep, "total cycles", cyc[k-1]-cyc[0], "total ns", ns[k-1]-ns[0], "=> MHz", (cyc[k-1]-cyc[0])/(ns[k-1]-ns[0])*1e3)
    for i in range(1, k):
        print(f"   {names[i]:14s} +{cyc[i]-cyc[i-1]:8d} cyc  +{ns[i]-ns[i-1]:8d} ns")
L = np.tril(dA.download((n, n)))
print("recon err", np.abs(L @ L.T - A).max())
