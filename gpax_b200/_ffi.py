"""
_ffi.py -- ctypes binding of libb200gp.so (include/b200gp.h).  The only module that touches the
C-ABI; everything numerical happens behind it on the GPU.  There is no CPU fallback: if the
library or a CUDA device is missing, importing works (so CPU-only hosts can run the host-logic
tests) but the first call raises `B200GPError`.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200gp.so")

KERNEL_RBF, KERNEL_MATERN52, KERNEL_PERIODIC = 0, 1, 2
KERNEL_NNGP_ERF, KERNEL_NNGP_RELU = 3, 4
KIND = {"RBF": KERNEL_RBF, "Matern": KERNEL_MATERN52, "Periodic": KERNEL_PERIODIC}

FLAG_DEVICE_PTRS = 1 << 0
FLAG_LOWER_ONLY = 1 << 1
FLAG_F32 = 1 << 2
OUT_MEAN, OUT_VAR, OUT_COV, OUT_SAMPLE = 1 << 4, 1 << 5, 1 << 6, 1 << 7


class B200GPError(RuntimeError):
    pass


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("gram_ms", C.c_double), ("potrf_ms", C.c_double),
                ("trsm_ms", C.c_double), ("epilogue_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("flops", C.c_double), ("gram_bytes", C.c_double),
                ("launches", C.c_int64), ("host_enqueue_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# every symbol include/b200gp.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "b2gp_version": (C.c_int, []),
    "b2gp_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "b2gp_ctx_destroy": (C.c_int, [_vp]),
    "b2gp_last_error": (C.c_char_p, [_vp]),
    "b2gp_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "b2gp_device_info": (C.c_int, [_vp, _ip, _ip, _ip, C.POINTER(C.c_size_t)]),
    "b2gp_last_timing": (C.c_int, [_vp, C.POINTER(Timing)]),
    "b2gp_dev_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "b2gp_dev_free": (C.c_int, [_vp, _vp]),
    "b2gp_host_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "b2gp_host_free": (C.c_int, [_vp, _vp]),
    "b2gp_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "b2gp_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "b2gp_sync": (C.c_int, [_vp]),
    "b2gp_gram": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, C.c_int, _vp, C.c_double, C.c_double,
                            C.c_double, C.c_int, _vp, C.c_int64, C.c_uint]),
    "b2gp_gram_multitask": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int, _vp, C.c_double, C.c_double, _vp,
                                      C.c_int, _vp, C.c_double, C.c_int, C.c_int, _vp, C.c_int64, C.c_uint]),
    "b2gp_potrf": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _ip, C.c_uint]),
    "b2gp_trsm_lower": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, C.c_uint]),
    "b2gp_gemm_nt": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_double, _vp, C.c_int64, _vp, C.c_int64,
                               C.c_double, _vp, C.c_int64, C.c_int, C.c_uint]),
    "b2gp_posterior": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, C.c_int, C.c_int64,
                                 _vp, C.c_int, C.c_double, C.c_uint, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp,
                                 C.POINTER(Timing)]),
    "b2gp_posterior_batch": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int,
                                       C.c_int64, _vp, _vp, C.c_int64, C.c_int, C.c_double, C.c_uint, _vp, _vp, _vp, _vp,
                                       C.c_int64, _vp, _vp, C.POINTER(Timing)]),
    "b2gp_mll_v": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int, _vp, _vp, C.c_double, C.c_uint, _dp, _vp, _vp, _vp, _ip]),
    "b2gp_sparse_posterior": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int,
                                        _vp, C.c_int, C.c_double, C.c_uint, _vp, _vp, _vp, _vp, C.POINTER(Timing)]),
    "b2gp_mll": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int, _vp, C.c_double, C.c_uint, _dp, _vp, _vp, _ip]),
    "b2gp_sparse_elbo": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int, _vp, C.c_double, C.c_uint, _dp, _vp,
                                   _vp, _ip]),
    "b2gp_dist_unique_id": (C.c_int, [_vp]),
    "b2gp_dist_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b2gp_dist_info": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "b2gp_dist_finalize": (C.c_int, [_vp]),
    "b2gp_dist_posterior": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int, _vp, C.c_int, C.c_double,
                                      C.c_int64, C.c_uint, _vp, _vp, _ip, C.POINTER(Timing)]),
    "b2gp_dist_sparse_posterior": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int, _vp, C.c_int,
                                             C.c_double, C.c_uint, _vp, _vp, _ip, C.POINTER(Timing)]),
    "b2gp_dist_layout": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                   C.POINTER(C.c_int64)]),
    "b2gp_sparse_partial": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int, _vp, C.c_double, _vp,
                                      C.c_int64, _vp, _ip]),
    "b2gp_sparse_finish": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int, _vp, C.c_int,
                                     C.c_double, C.c_uint, _vp, _vp, _vp, _ip]),
    "b2gp_potrf_inv": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, _ip]),
    "b2gp_trsm_inv": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, _vp, C.c_int64]),
    "b2gp_rowdot": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, C.c_double, _vp, _vp, C.c_int]),
    "b2gp_mvn_sample": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, _vp, C.c_uint]),
    "b2gp_acq_moments": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int, _vp,
                                   C.c_uint]),
    "b2gp_acq_samples": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int, _vp, _vp,
                                   _vp, C.c_uint]),
    "b2gp_kg": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp, C.c_int64, C.c_double, C.c_double, C.c_int, _vp, C.c_uint]),
    "b2gp_copy2d": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libb200gp.so and bind every declared symbol (no CUDA call is made)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200GPError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C gpax_b200/csrc`).  gpax_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class DeviceArray:
    """A device allocation owned by a Context (fp64 elements unless stated)."""

    def __init__(self, ctx, nbytes, shape=None):
        self.ctx, self.nbytes, self.shape = ctx, int(nbytes), shape
        p = C.c_void_p()
        ctx._check(ctx.lib.b2gp_dev_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p

    def upload(self, host):
        host = np.ascontiguousarray(host)
        assert host.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.b2gp_h2d(self.ctx.h, self.ptr, _ptr(host), host.nbytes))
        return self

    def download(self, shape=None, dtype=np.float64):
        shape = shape if shape is not None else self.shape
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        self.ctx._check(self.ctx.lib.b2gp_d2h(self.ctx.h, _ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr is not None and self.ctx.h is not None:
            self.ctx.lib.b2gp_dev_free(self.ctx.h, self.ptr)
        self.ptr = None


class Context:
    """One per Python thread (a ctx is not thread-safe).  Owns streams, workspaces, the device."""

    def __init__(self, device=0, streams=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.b2gp_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise B200GPError(f"b2gp_ctx_create(device={device}) failed with {rc}: no usable CUDA device "
                              "(gpax_b200 has no CPU fallback)")
        self.h = h
        self.device = device
        self._pinned = []          # pinned host blocks handed out by pinned(): released with the context
        if streams is not None:
            self.set_option("streams", streams)

    # ---- plumbing
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.b2gp_last_error(self.h)
            raise B200GPError(f"libb200gp error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        """destroy the context; pinned arrays obtained from pinned() must not be used afterwards"""
        if getattr(self, "h", None) is not None:
            for p in getattr(self, "_pinned", []):
                self.lib.b2gp_host_free(self.h, p)
            self._pinned = []
            self.lib.b2gp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._check(self.lib.b2gp_set_option(self.h, key.encode(), int(value)))

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self.lib.b2gp_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "mem_bytes": mem.value}

    def last_timing(self):
        t = Timing()
        self._check(self.lib.b2gp_last_timing(self.h, C.byref(t)))
        return t.as_dict()

    def sync(self):
        self._check(self.lib.b2gp_sync(self.h))

    def alloc(self, shape, dtype=np.float64):
        shape = tuple(np.atleast_1d(shape).tolist())
        return DeviceArray(self, int(np.prod(shape)) * np.dtype(dtype).itemsize, shape)

    def to_device(self, host):
        host = np.ascontiguousarray(host)
        return DeviceArray(self, host.nbytes, host.shape).upload(host)

    def pinned(self, shape, dtype=np.float64):
        """NumPy array over pinned (page-locked) host memory; the block lives until close()"""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._check(self.lib.b2gp_host_alloc(self.h, n, C.byref(p)))
        self._pinned.append(C.c_void_p(p.value))
        buf = (C.c_char * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        return arr

    # ---- primitives (host arrays in / out)
    def gram(self, kind, X, Z, lengthscale, scale, period=1.0, diag_add=0.0, same_xz=None, lower_only=False, f32=False):
        """f32: X, Z are handed over as float32 and K comes back float32 (B2GP_FLAG_F32)"""
        cv, dt = (_f32, np.float32) if f32 else (_f64, np.float64)
        X, Z = cv(X), cv(Z)
        n, d = X.shape
        m = Z.shape[0]
        ell = _f64(np.broadcast_to(np.asarray(lengthscale, dtype=np.float64).reshape(-1), (d,)))
        if same_xz is None:
            same_xz = X.shape == Z.shape          # kernels.py:63
        K = np.zeros((n, m), dtype=dt) if lower_only else np.empty((n, m), dtype=dt)
        if n == 0 or m == 0:
            return K
        flags = (FLAG_LOWER_ONLY if lower_only else 0) | (FLAG_F32 if f32 else 0)
        self._check(self.lib.b2gp_gram(self.h, KIND[kind] if isinstance(kind, str) else kind, _ptr(X), n, _ptr(Z), m, d,
                                       _ptr(ell), float(scale), float(period), float(diag_add), int(bool(same_xz)),
                                       _ptr(K), m, flags))
        return K

    def gram_multitask(self, kind, X, tX, Z, tZ, lengthscale, scale, period, B, noise_task, jitter, same_xz, group=1):
        X, Z, B = _f64(X), _f64(Z), _f64(B)
        n, d = X.shape
        m = Z.shape[0]
        tX, tZ = np.ascontiguousarray(tX, dtype=np.int32), np.ascontiguousarray(tZ, dtype=np.int32)
        ell = _f64(np.broadcast_to(np.asarray(lengthscale, dtype=np.float64).reshape(-1), (d,)))
        nt = None if noise_task is None else _f64(noise_task)
        K = np.empty((n, m))
        self._check(self.lib.b2gp_gram_multitask(self.h, kind if isinstance(kind, int) else KIND[kind], _ptr(X), _ptr(tX), n, _ptr(Z), _ptr(tZ),
                                                 m, d, _ptr(ell), float(scale), float(period), _ptr(B), B.shape[0], _ptr(nt), float(jitter),
                                                 int(bool(same_xz)), int(group), _ptr(K), m, 0))
        return K

    def potrf(self, A):
        A = np.array(A, dtype=np.float64, order="C", copy=True)
        n = A.shape[0]
        info = C.c_int(0)
        self._check(self.lib.b2gp_potrf(self.h, n, _ptr(A), A.shape[1] if A.ndim == 2 else 1, C.byref(info), 0))
        return A, info.value

    def trsm_lower(self, L, B):
        """Rows of B are right-hand sides: returns B L^{-T} (row r = L^{-1} b_r).  Call right after potrf."""
        L = _f64(L)
        B = np.array(B, dtype=np.float64, order="C", copy=True)
        n = L.shape[0]
        self._check(self.lib.b2gp_trsm_lower(self.h, n, B.shape[0], _ptr(L), n, _ptr(B), B.shape[1], 0))
        return B

    def gemm_nt(self, A, B, C_in=None, alpha=1.0, beta=0.0, lower_only=False):
        A, B = _f64(A), _f64(B)
        m, k = A.shape
        n = B.shape[0]
        Cm = np.zeros((m, n)) if C_in is None else np.array(C_in, dtype=np.float64, order="C", copy=True)
        self._check(self.lib.b2gp_gemm_nt(self.h, m, n, k, float(alpha), _ptr(A), k, _ptr(B), k, float(beta), _ptr(Cm), n,
                                          int(lower_only), 0))
        return Cm

    # ---- the posterior (host arrays in / out)
    def posterior(self, kind, Xtr, yres, Xnew, theta, noiseless=False, jitter=1e-6, want=("mean", "cov"),
                  eps=None, timing=False, noise_vec=None, f32=False):
        """theta: [S, d+3] rows (lengthscale[d], k_scale, noise, period); yres [N] or [S, N].  Xtr [N, d] or per member
        [S, N, d]; Xnew [P, d] or [S, P, d]; noise_vec None, [N] or [S, N] (per-point noise variances on k_XX's diagonal).
        f32: the data arrays cross the boundary as float32 in both directions (B2GP_FLAG_F32), theta stays float64."""
        _f64 = _f32 if f32 else globals()["_f64"]      # noqa: F811  data arrays in the I/O precision
        odt = np.float32 if f32 else np.float64
        Xtr, Xnew = _f64(Xtr), _f64(Xnew)
        N, d = Xtr.shape[-2:]
        P = Xnew.shape[-2]
        theta = globals()["_f64"](theta).reshape(-1, d + 3)
        S = theta.shape[0]
        yres = _f64(yres)
        stride = 0 if yres.ndim == 1 else yres.shape[1]
        xs = 0 if Xtr.ndim == 2 else N * d
        xns = 0 if Xnew.ndim == 2 else P * d
        nv = None if noise_vec is None else _f64(noise_vec)
        nvs = 0 if nv is None or nv.ndim == 1 else N
        for a, n_ in ((Xtr, xs), (Xnew, xns), (nv, nvs)):
            if a is not None and n_ and a.shape[0] != S:
                raise ValueError("per-member arrays need a leading axis of length S = theta.shape[0]")
        flags = FLAG_F32 if f32 else 0
        mean = var = cov = samp = None
        if "mean" in want:
            flags |= OUT_MEAN
            mean = np.empty((S, P), dtype=odt)
        if "var" in want:
            flags |= OUT_VAR
            var = np.empty((S, P), dtype=odt)
        if "cov" in want:
            flags |= OUT_COV
            cov = np.empty((S, P, P), dtype=odt)
        n_samp = 0
        if eps is not None:
            eps = _f64(eps).reshape(S, -1, P)
            n_samp = eps.shape[1]
            flags |= OUT_SAMPLE
            samp = np.empty((S, n_samp, P), dtype=odt)
        info = np.zeros(S, dtype=np.int32)
        t = Timing()
        self._check(self.lib.b2gp_posterior_batch(
            self.h, KIND[kind] if isinstance(kind, str) else kind, _ptr(Xtr), xs, N, _ptr(yres), stride, _ptr(Xnew), xns, P, d, S,
            _ptr(theta), _ptr(nv), nvs, int(bool(noiseless)), float(jitter), flags, _ptr(mean), _ptr(var), _ptr(cov), _ptr(eps),
            n_samp, _ptr(samp), info.ctypes.data_as(_vp), C.byref(t) if timing else None))
        out = {"mean": mean, "var": var, "cov": cov, "y_sampled": samp, "info": info}
        if timing:
            out["timing"] = t.as_dict()
        return out

    def mll(self, kind, X, yres, theta, jitter=1e-6, want_grad=True, want_alpha=False, noise_vec=None):
        """log marginal likelihood, its gradient w.r.t. log(lengthscale[d], k_scale, noise, period), alpha = K^-1 y.
        With noise_vec [N] (per-point noise variances on the diagonal) the return gains d value / d noise_vec."""
        X, yres = _f64(X), _f64(yres)
        N, d = X.shape
        theta = _f64(theta).reshape(d + 3)
        val = C.c_double(0.0)
        grad = np.zeros(d + 3) if want_grad else None
        alpha = np.zeros(N) if want_alpha else None
        info = C.c_int(0)
        if noise_vec is None:
            self._check(self.lib.b2gp_mll(self.h, KIND[kind] if isinstance(kind, str) else kind, _ptr(X), N, _ptr(yres), d,
                                          _ptr(theta), float(jitter), 0, C.byref(val), _ptr(grad), _ptr(alpha), C.byref(info)))
            return val.value, grad, alpha, info.value
        nv = _f64(noise_vec).reshape(N)
        gnv = np.zeros(N) if want_grad else None
        self._check(self.lib.b2gp_mll_v(self.h, KIND[kind] if isinstance(kind, str) else kind, _ptr(X), N, _ptr(yres), d,
                                        _ptr(theta), _ptr(nv), float(jitter), 0, C.byref(val), _ptr(grad), _ptr(alpha), _ptr(gnv),
                                        C.byref(info)))
        return val.value, grad, alpha, info.value, gnv

    def sparse_elbo(self, kind, Xu, X, yres, theta, jitter=1e-6):
        """VFE bound of the sparse GP, its gradient w.r.t. log(lengthscale[d], k_scale, noise, period) and w.r.t. Xu"""
        Xu, X, yres = _f64(Xu), _f64(X), _f64(yres)
        M, d = Xu.shape
        N = X.shape[0]
        theta = _f64(theta).reshape(d + 3)
        val, info = C.c_double(0.0), C.c_int(0)
        g, gx = np.zeros(d + 3), np.zeros((M, d))
        self._check(self.lib.b2gp_sparse_elbo(self.h, KIND[kind] if isinstance(kind, str) else kind, _ptr(Xu), M, _ptr(X), N,
                                              _ptr(yres), d, _ptr(theta), float(jitter), 0, C.byref(val), _ptr(g), _ptr(gx),
                                              C.byref(info)))
        return val.value, g, gx, info.value

    def sparse_posterior(self, kind, Xu, Xtr, yres, Xnew, theta, noiseless=False, jitter=1e-6, want=("mean", "cov")):
        Xu, Xtr, Xnew = _f64(Xu), _f64(Xtr), _f64(Xnew)
        M, d = Xu.shape
        N, P = Xtr.shape[0], Xnew.shape[0]
        theta = _f64(theta).reshape(d + 3)
        yres = _f64(yres).reshape(N)
        flags = 0
        mean = var = cov = None
        if "mean" in want:
            flags |= OUT_MEAN
            mean = np.empty(P)
        if "var" in want:
            flags |= OUT_VAR
            var = np.empty(P)
        if "cov" in want:
            flags |= OUT_COV
            cov = np.empty((P, P))
        info = np.zeros(1, dtype=np.int32)
        t = Timing()
        self._check(self.lib.b2gp_sparse_posterior(
            self.h, KIND[kind] if isinstance(kind, str) else kind, _ptr(Xu), M, _ptr(Xtr), N, _ptr(yres), _ptr(Xnew), P, d,
            _ptr(theta), int(bool(noiseless)), float(jitter), flags, _ptr(mean), _ptr(var), _ptr(cov),
            info.ctypes.data_as(_vp), C.byref(t)))
        return {"mean": mean, "var": var, "cov": cov, "info": int(info[0]), "timing": t.as_dict()}


ACQ = {"EI": 0, "UCB": 1, "UE": 2, "POI": 3}


def _acq_moments(self, kind, mean, var, best_f=None, param=0.0, maximize=False):
    """acquisition function `kind` on moments [P] or [R, P] (host arrays)"""
    var = _f64(var)
    shape = var.shape
    var = var.reshape(-1, shape[-1])
    mean = None if mean is None else _f64(mean).reshape(var.shape)
    out = np.empty_like(var)
    self._check(self.lib.b2gp_acq_moments(self.h, ACQ[kind], _ptr(mean), _ptr(var), var.shape[0], var.shape[1],
                                          int(best_f is not None), float(best_f or 0.0), float(param), int(bool(maximize)),
                                          _ptr(out), 0))
    return out.reshape(shape)


def _acq_samples(self, kind, y, best_f=None, param=0.0, maximize=False):
    """moments over the rows of y [R, P], then the acquisition function; returns (acq, mean, var)"""
    y = _f64(y)
    y = y.reshape(-1, y.shape[-1])
    R, P = y.shape
    out, mean, var = np.empty(P), np.empty(P), np.empty(P)
    self._check(self.lib.b2gp_acq_samples(self.h, ACQ[kind], _ptr(y), R, P, int(best_f is not None), float(best_f or 0.0),
                                          float(param), int(bool(maximize)), _ptr(out), _ptr(mean), _ptr(var), 0))
    return out, mean, var


def _kg(self, mean, cov, ysim, diag_sub, noise_plus_jitter, maximize=True):
    mean, cov, ysim = _f64(mean), _f64(cov), _f64(ysim)
    P = mean.shape[0]
    ysim = ysim.reshape(-1, P)
    out = np.empty(P)
    self._check(self.lib.b2gp_kg(self.h, _ptr(mean), _ptr(cov), P, _ptr(ysim), ysim.shape[0], float(diag_sub),
                                 float(noise_plus_jitter), int(bool(maximize)), _ptr(out), 0))
    return out


def _mvn_sample(self, mean, cov, eps):
    """mean [S, P], cov [S, P, P], eps [S, n, P] -> (mean + chol(cov) eps [S, n, P], info [S])"""
    mean, cov, eps = _f64(mean), _f64(cov), _f64(eps)
    P = mean.shape[-1]
    mean, cov = mean.reshape(-1, P), cov.reshape(-1, P, P)
    S = mean.shape[0]
    eps = eps.reshape(S, -1, P)
    y = np.empty_like(eps)
    info = np.zeros(S, dtype=np.int32)
    self._check(self.lib.b2gp_mvn_sample(self.h, _ptr(mean), _ptr(cov), S, P, _ptr(eps), eps.shape[1], _ptr(y),
                                         info.ctypes.data_as(_vp), 0))
    return y, info


Context.acq_moments, Context.acq_samples, Context.kg, Context.mvn_sample = _acq_moments, _acq_samples, _kg, _mvn_sample

_default_ctx = None


def default_context():
    """Process-wide context on the device named by LOCAL_RANK / B200GP_DEVICE (default 0)."""
    global _default_ctx
    if _default_ctx is None or _default_ctx.h is None:
        dev = int(os.environ.get("B200GP_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _default_ctx = Context(dev)
    return _default_ctx
