"""ozaki_probe.py -- int8-tcgen05 (Ozaki) GEMM/SYRK vs the DMMA kernel and NumPy; timing at the headline SYRK size."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from gpax_b200 import _ffi  # noqa: E402

ctx = _ffi.Context(0)
fn = ctx.lib.b2gp_debug_ozaki
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int] + [C.c_int64] * 3 + [C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                        C.c_int, C.POINTER(C.c_double)]
rng = np.random.default_rng(0)


def run(S, m, n, k, lower, same, check=True, reps=1):
    A = rng.standard_normal((m, k)) * np.exp(rng.normal(0, 2, (m, 1)))      # rows of very different scale
    B = A if same else rng.standard_normal((n, k)) * np.exp(rng.normal(0, 2, (n, 1)))
    C0 = rng.standard_normal((m, n))
    dA = ctx.to_device(A)
    dB = dA if same else ctx.to_device(B)
    dC = ctx.to_device(C0)
    ms = C.c_double()
    best = 1e9
    for r in range(reps):
        dC.upload(C0)
        rc = fn(ctx.h, S, m, n, k, -1.0, dA.ptr, k, dB.ptr, k, dC.ptr, n, int(lower), C.byref(ms))
        assert rc == 0, ctx.lib.b2gp_last_error(ctx.h)
        best = min(best, ms.value)
    out = dC.download((m, n))
    msg = f"S={S} m={m} n={n} k={k} lower={int(lower)}: {best:8.3f} ms  {(1 if lower else 2) * m * n * k / best / 1e9:7.2f} TF/s-eq"
    if check:
        ref = C0 - A @ B.T
        scale = np.abs(A).sum(1)[:, None] * 0 + (np.linalg.norm(A, axis=1)[:, None] * np.linalg.norm(B, axis=1)[None, :])
        if lower:
            mask = np.tril(np.ones((m, n), bool))
            err = (np.abs(out - ref) / scale)[mask].max()
            assert np.array_equal(out[~mask], C0[~mask]), "upper triangle touched"
        else:
            err = (np.abs(out - ref) / scale).max()
        msg += f"   max |err| / (|a_i||b_j|) = {err:.2e}"
    print(msg, flush=True)
    for d in {id(dA): dA, id(dB): dB, id(dC): dC}.values():
        d.free()


mode = sys.argv[1] if len(sys.argv) > 1 else "small"
if mode == "small":
    for S in (7, 6):
        run(S, 128, 64, 64, False, False)
        run(S, 128, 64, 256, False, False)
        run(S, 256, 192, 128, False, False)
        run(S, 300, 100, 70, False, False)
        run(S, 512, 512, 200, True, True)
        run(S, 1024, 1024, 1024, True, True)
elif mode == "prof":          # run with B2GP_OZ_PROF=1: issuer / epilogue cycle counters per tile; cluster 2 and 1
    for cl in (2, 1):
        ctx.set_option("oz_cluster", cl)
        print(f"--- oz_cluster = {cl}", flush=True)
        for S in (6, 7):
            for kk in (512, 1024, 8192):
                run(S, 8192, 8192, kk, True, True, check=False, reps=2)
            run(S, 8192, 8192, 2048, False, False, check=False, reps=2)
    ctx.set_option("oz_cluster", 2)
else:
    for S in (7, 6):
        run(S, 4096, 4096, 4096, True, True, check=True, reps=2)
        run(S, 8192, 8192, 8192, True, True, check=False, reps=3)
        run(S, 8192, 8192, 8192, False, False, check=False, reps=2)
        run(S, 8192, 8192, 512, True, True, check=False, reps=3)
