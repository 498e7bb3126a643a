"""
kernels.py -- the reference's kernel seam (gpax/kernels/kernels.py:17):
    K = kernel(X, Z, params, noise=0, jitter=1e-6)
for RBFKernel (kernels.py:44-65), MaternKernel (68-91), PeriodicKernel (94-117) and the name table
get_kernel (227-241).  Each call is one fused Gram build on the GPU (b2gp_gram); the arrays that come
back are NumPy fp64.
"""
from typing import Callable, Dict, Union

import numpy as np

from . import _ffi

kernel_fn_type = Callable[[np.ndarray, np.ndarray, Dict[str, np.ndarray], np.ndarray], np.ndarray]


def _as2d(X):
    X = np.asarray(X)
    if X.dtype != np.float32:                       # float32 inputs keep their precision across the boundary (B2GP_FLAG_F32)
        X = X.astype(np.float64, copy=False)
    return X if X.ndim > 1 else X[:, None]


def _scalar(v, name):
    a = np.asarray(v, dtype=np.float64)
    if a.size != 1:
        raise ValueError(f"'{name}' must be a scalar for this kernel, got shape {a.shape}")
    return float(a.reshape(()))


def _gram(kind, X, Z, params, noise, jitter, ctx=None):
    X, Z = _as2d(X), _as2d(Z)
    if X.shape[1] != Z.shape[1]:
        raise ValueError(f"feature dimensions differ: {X.shape} vs {Z.shape}")
    ctx = ctx or _ffi.default_context()
    period = params.get("period", None)
    period = 1.0 if period is None else _scalar(period, "period")
    same = X.shape == Z.shape                       # kernels.py:63 -- shape equality, not identity
    diag = _scalar(noise, "noise") + float(jitter)  # kernels.py:24-25, 64
    f32 = np.asarray(X).dtype == np.float32 and np.asarray(Z).dtype == np.float32      # the reference's default precision
    return ctx.gram(kind, X, Z, params["k_length"], _scalar(params["k_scale"], "k_scale"), period, diag, same, f32=f32)


def RBFKernel(X, Z, params, noise=0, jitter=1e-6, **kwargs):
    """Radial basis function kernel (gpax/kernels/kernels.py:44-65)."""
    return _gram("RBF", X, Z, params, noise, jitter, kwargs.get("ctx"))


def MaternKernel(X, Z, params, noise=0, jitter=1e-6, **kwargs):
    """Matern-5/2 kernel (gpax/kernels/kernels.py:68-91)."""
    return _gram("Matern", X, Z, params, noise, jitter, kwargs.get("ctx"))


def PeriodicKernel(X, Z, params, noise=0, jitter=1e-6, **kwargs):
    """Periodic kernel (gpax/kernels/kernels.py:94-117); params carries 'period'."""
    return _gram("Periodic", X, Z, params, noise, jitter, kwargs.get("ctx"))


_BOOK = {"RBF": RBFKernel, "Matern": MaternKernel, "Periodic": PeriodicKernel}
_NAME_OF = {RBFKernel: "RBF", MaternKernel: "Matern", PeriodicKernel: "Periodic"}


def NNGPKernel(activation: str = "erf", depth: int = 3):
    """Neural-network GP kernel (gpax/kernels/kernels.py:186-224): returns k(X, Z, params, noise, jitter) with
    params = {"var_b", "var_w"}; the recursion over `depth` layers runs in the fused Gram kernel on the GPU."""
    kind = _ffi.KERNEL_NNGP_RELU if activation == "relu" else _ffi.KERNEL_NNGP_ERF

    def NNGPKernel_func(X, Z, params, noise=0, jitter=1e-6, **kwargs):
        X, Z = _as2d(X), _as2d(Z)
        ctx = kwargs.get("ctx") or _ffi.default_context()
        same = X.shape == Z.shape                                    # kernels.py:221
        f32 = X.dtype == np.float32 and Z.dtype == np.float32
        ell = np.full(X.shape[1], float(depth))                      # lengthscale slot 0 carries the depth
        return ctx.gram(kind, X, Z, ell, _scalar(params["var_w"], "var_w"), _scalar(params["var_b"], "var_b"),
                        _scalar(noise, "noise") + float(jitter), same, f32=f32)
    return NNGPKernel_func


def get_kernel(kernel: Union[str, kernel_fn_type] = "RBF", **kwargs):
    """Name -> function; callables pass through (gpax/kernels/kernels.py:225-241).  'NNGP' takes activation= / depth=
    through **kwargs as in the reference; it is a Gram-only kernel here (non-stationary), so models built on it go
    through the callable-kernel posterior (host-supplied Gram matrices, GPU factorisation and solves)."""
    if isinstance(kernel, str):
        if kernel == "NNGP":
            return NNGPKernel(**kwargs)
        try:
            kernel = _BOOK[kernel]
        except KeyError:
            print("Select one of the currently available kernels:", *_BOOK.keys(), "NNGP")
            raise
    return kernel


def builtin_name(kernel):
    """'RBF' / 'Matern' / 'Periodic' when `kernel` is (or names) one of the fused GPU kernels, else None."""
    if isinstance(kernel, str):
        return kernel if kernel in _BOOK else None
    return _NAME_OF.get(kernel)
