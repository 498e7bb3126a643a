"""perf_probe.py -- quick device-resident timings of the kernels (development tool, not the bench)."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gpax_b200 import _ffi  # noqa: E402


def main():
    ctx = _ffi.Context(0)
    lib, h = ctx.lib, ctx.h
    out = {"device": ctx.device_info()}
    try:
        import torch
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        for _ in range(2):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["cublas_dgemm_8192_tflops"] = 2 * 8192 ** 3 / best / 1e9
        del a, b
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["cublas_dgemm_error"] = repr(e)

    rng = np.random.default_rng(0)
    for n in (4096, 8192):
        A = ctx.to_device(rng.standard_normal((n, n)))
        Cm = ctx.to_device(np.zeros((n, n)))
        for lower in (0, 1):
            best = 1e9
            for _ in range(3):
                ctx._check(lib.b2gp_gemm_nt(h, n, n, n, -1.0, A.ptr, n, A.ptr, n, 1.0, Cm.ptr, n, lower, _ffi.FLAG_DEVICE_PTRS))
                best = min(best, ctx.last_timing()["epilogue_ms"])
            fl = (1 if lower else 2) * n ** 3
            out[f"gemm_n{n}_lower{lower}"] = {"ms": best, "tflops": fl / best / 1e9}
        A.free()
        Cm.free()

    for N in (2048, 8192, 16384):
        d = 3
        X = ctx.to_device(rng.uniform(0, 1, (N, d)))
        ell = np.full(d, 0.3)
        K = ctx.alloc((N, N))
        res = {}
        for rep in range(2):
            ctx._check(lib.b2gp_gram(h, 0, X.ptr, N, X.ptr, N, d, _ffi._ptr(ell), 1.0, 1.0, 0.1 + 1e-6, 1, K.ptr, N,
                                     _ffi.FLAG_DEVICE_PTRS))
            t = ctx.last_timing()
            res["gram_full_ms"] = t["total_ms"]
            ctx._check(lib.b2gp_gram(h, 0, X.ptr, N, X.ptr, N, d, _ffi._ptr(ell), 1.0, 1.0, 0.1 + 1e-6, 1, K.ptr, N,
                                     _ffi.FLAG_DEVICE_PTRS | _ffi.FLAG_LOWER_ONLY))
            res["gram_lower_ms"] = ctx.last_timing()["total_ms"]
            info = C.c_int(0)
            t0 = time.perf_counter()
            ctx._check(lib.b2gp_potrf(h, N, K.ptr, N, C.byref(info), _ffi.FLAG_DEVICE_PTRS))
            res["potrf_wall_ms"] = (time.perf_counter() - t0) * 1e3
            t = ctx.last_timing()
            res["potrf_ms"] = t["total_ms"]
            res["potrf_tflops"] = N ** 3 / 3 / t["total_ms"] / 1e9
            res["potrf_launches"] = t["launches"]
            res["info"] = info.value
        res["gram_full_gbs"] = 8 * N * N / res["gram_full_ms"] / 1e6
        out[f"N{N}"] = res
        X.free()
        K.free()

    # full posterior, host pointers, S draws
    N, P, d, S = 16384, 1024, 3, 4
    Xtr = rng.uniform(0, 1, (N, d))
    y = np.sin(3 * Xtr[:, 0]) + 0.1 * rng.standard_normal(N)
    Xn = rng.uniform(0, 1, (P, d))
    theta = np.tile(np.array([0.3, 0.3, 0.3, 1.0, 0.1, 1.0]), (S, 1))
    for streams in (1, 2, 3):
        ctx.set_option("streams", streams)
        for rep in range(2):
            t0 = time.perf_counter()
            o = ctx.posterior("RBF", Xtr, y, Xn, theta, want=("mean", "var"), timing=True)
            wall = time.perf_counter() - t0
        o["timing"]["wall_ms"] = wall * 1e3
        o["timing"]["posteriors_per_s"] = S / wall
        o["timing"]["info"] = o["info"].tolist()
        out[f"posterior_N{N}_P{P}_S{S}_streams{streams}"] = o["timing"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
