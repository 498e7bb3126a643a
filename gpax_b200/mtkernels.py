"""
mtkernels.py -- multi-task kernels with the reference's surface (gpax/kernels/mtkernels.py): `index_kernel` (:19-58),
`MultitaskKernel` (:61-125), `MultivariateKernel` (:128-192), `LCMKernel` (:195-232).  The data kernel and the
task-covariance factor are evaluated on the GPU in one C-ABI call (b2gp_gram_multitask); the callables returned here
plug into `ExactGP(kernel=...)` like any user kernel (Gram matrices from the callable, factorisation and solves on the GPU).
"""
import numpy as np

from . import _ffi
from .kernels import _as2d, _scalar

_KIND = {"RBF": _ffi.KERNEL_RBF, "Matern": _ffi.KERNEL_MATERN52, "Periodic": _ffi.KERNEL_PERIODIC}


def index_kernel(indices1, indices2, params):
    """mtkernels.py:19-58: B[indices1, indices2] with B = W W^T + diag(v)."""
    W, v = np.asarray(params["W"], dtype=np.float64), np.asarray(params["v"], dtype=np.float64)
    B = W @ W.T + np.diag(v)
    return B[np.ix_(np.asarray(indices1, dtype=int), np.asarray(indices2, dtype=int))]


def _base(base_kernel, kwargs1):
    """(kind, lengthscale(params, d), scale(params), period(params)) of the fused data kernel"""
    if base_kernel in _KIND:
        kind = _KIND[base_kernel]
        return kind, (lambda p, d: p["k_length"]), (lambda p: _scalar(p["k_scale"], "k_scale")), \
            (lambda p: 1.0 if p.get("period") is None else _scalar(p["period"], "period"))
    if base_kernel == "NNGP":
        kind = _ffi.KERNEL_NNGP_RELU if kwargs1.get("activation", "erf") == "relu" else _ffi.KERNEL_NNGP_ERF
        depth = float(kwargs1.get("depth", 3))
        return kind, (lambda p, d: np.full(d, depth)), (lambda p: _scalar(p["var_w"], "var_w")), (lambda p: _scalar(p["var_b"], "var_b"))
    raise NotImplementedError("multi-task kernels are fused for the built-in data kernels 'RBF', 'Matern', 'Periodic', 'NNGP'")


def _task_matrix(params):
    W, v = np.asarray(params["W"], dtype=np.float64), np.asarray(params["v"], dtype=np.float64)
    return W @ W.T + np.diag(v)                                   # mtkernels.py:55-57


def MultitaskKernel(base_kernel, **kwargs1):
    """mtkernels.py:61-125: K(x_i, y_j) = k_data(x, y) * k_task(i, j); task indices in the last input column."""
    kind, ell_of, scale_of, period_of = _base(base_kernel, kwargs1)

    def multi_task_kernel(X, Z, params, noise=0, **kwargs2):
        X, Z = _as2d(X).astype(np.float64), _as2d(Z).astype(np.float64)
        ctx = kwargs2.get("ctx") or _ffi.default_context()
        Xd, tX = X[:, :-1], X[:, -1].astype(int)
        Zd, tZ = Z[:, :-1], Z[:, -1].astype(int)
        B = _task_matrix(params)
        same = X.shape == Z.shape                                 # mtkernels.py:111
        nt = np.ones(1) * noise if isinstance(noise, (int, float)) else np.asarray(noise, dtype=np.float64)   # :113-114
        nt = np.broadcast_to(nt, (B.shape[0],)) if nt.size == 1 else nt
        return ctx.gram_multitask(kind, Xd, tX, Zd, tZ, ell_of(params, Xd.shape[1]), scale_of(params), period_of(params), B, nt,
                                  float(kwargs2.get("jitter", 1e-6)), same)
    return multi_task_kernel


def MultivariateKernel(base_kernel, num_tasks, **kwargs1):
    """mtkernels.py:128-192: K = kron(k_data, k_task) for tasks sharing one input space -- the multi-task kernel on the
    inputs repeated once per task with the task index cycling fastest."""
    kind, ell_of, scale_of, period_of = _base(base_kernel, kwargs1)

    def multivariate_kernel(X, Z, params, noise=0, **kwargs2):
        X, Z = _as2d(X).astype(np.float64), _as2d(Z).astype(np.float64)
        ctx = kwargs2.get("ctx") or _ffi.default_context()
        T = int(num_tasks)
        B = _task_matrix(params)
        same = X.shape == Z.shape                                 # mtkernels.py:177
        nt = np.ones(T) * noise if isinstance(noise, (int, float)) else np.asarray(noise, dtype=np.float64)    # :179-180
        Xr, Zr = np.repeat(X, T, axis=0), np.repeat(Z, T, axis=0)
        tX, tZ = np.tile(np.arange(T), X.shape[0]), np.tile(np.arange(T), Z.shape[0])
        return ctx.gram_multitask(kind, Xr, tX, Zr, tZ, ell_of(params, X.shape[1]), scale_of(params), period_of(params), B, nt,
                                  float(kwargs2.get("jitter", 1e-6)), same, group=T)
    return multivariate_kernel


def LCMKernel(base_kernel, shared_input_space=True, num_tasks=None, **kwargs1):
    """mtkernels.py:195-232: sum over latent functions (leading axis of every parameter except the noise)."""
    multi = MultivariateKernel(base_kernel, num_tasks, **kwargs1) if shared_input_space else MultitaskKernel(base_kernel, **kwargs1)

    def lcm_kernel(X, Z, params, noise=0, **kwargs2):
        L = len(np.asarray(next(v for k, v in params.items() if k != "noise")))
        out = None
        for q in range(L):
            pq = {k: (v if k == "noise" else np.asarray(v)[q]) for k, v in params.items()}
            k = multi(X, Z, pq, noise, **kwargs2)
            out = k if out is None else out + k
        return out
    return lcm_kernel
