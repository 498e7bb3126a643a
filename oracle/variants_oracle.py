"""variants_oracle.py -- NumPy restatement of the model variants of SURVEY.md section 8f-3.  TEST INFRASTRUCTURE ONLY.

  var_noise_posterior     gpax/models/hskgp.py:165-206  (VarNoiseGP.get_mvn_posterior: main GP with zero noise + the
                          noise GP's predicted log-variance on the diagonal)
  task_batch_posterior    gpax/models/vgp.py:125-172    (vExactGP: vmap of the exact posterior over an outer task axis)
  uigp_posterior          gpax/models/uigp.py:131-150   (training inputs X_prime are a per-draw parameter)
  nngp_kernel             gpax/kernels/kernels.py:120-224
  measured_noise_logp     gpax/models/mngp.py:92-97     (log N(y; 0, k + diag(measured_noise)))
Pinned by tests/test_oracle_f.py against tests/golden/reference_vectors_f.npz."""
import numpy as np
import scipy.linalg as sla

from . import gp_oracle as go


def var_noise_posterior(X_train, y_train, X_new, params, kernel="RBF", noise_kernel="RBF", **kwargs):
    kern, nkern = go.get_kernel(kernel), go.get_kernel(noise_kernel)
    k_pp = kern(X_new, X_new, params, 0, **kwargs)                       # hskgp.py:176
    k_pX = kern(X_new, X_train, params, jitter=0.0)
    k_XX = kern(X_train, X_train, params, 0, **kwargs)
    Kinv = np.linalg.inv(k_XX)                                           # :180
    cov = k_pp - k_pX @ (Kinv @ k_pX.T)
    mean = k_pX @ (Kinv @ y_train)
    nparams = {"k_length": params["k_noise_length"], "k_scale": params["k_noise_scale"]}   # _set_noise_kernel_fn renames the sites
    k_pX_n = nkern(X_new, X_train, nparams, jitter=0.0)                  # :189
    k_XX_n = nkern(X_train, X_train, nparams, 0, **kwargs)
    pred_log_var = k_pX_n @ (np.linalg.inv(k_XX_n) @ params["log_var"])  # :196-197
    return mean, cov + np.diag(np.exp(pred_log_var))                     # :201-204


def task_batch_posterior(X_train, y_train, X_new, params, kernel="RBF", noiseless=False, **kwargs):
    means, covs = [], []
    for b in range(X_train.shape[0]):                                    # vgp.py:170-172 (vmap over the task axis)
        pb = {k: np.asarray(v)[b] for k, v in params.items()}
        m, c = go.exact_posterior(X_train[b], y_train[b], X_new[b], pb, kernel, noiseless, **kwargs)
        means.append(m)
        covs.append(c)
    return np.stack(means), np.stack(covs)


def uigp_posterior(y_train, X_new, params, kernel="RBF", noiseless=False, **kwargs):
    return go.exact_posterior(params["X_prime"], y_train, X_new, params, kernel, noiseless, **kwargs)   # uigp.py:138-150


def _nngp_pair(x1x2, x1x1, x2x2, var_b, var_w, depth, act, d):
    k12, k11, k22 = (var_b + var_w * v / d for v in (x1x2, x1x1, x2x2))      # depth 0, kernels.py:139-140
    for _ in range(depth):
        if act == "erf":
            f = lambda a, b, c: var_b + 2 * var_w / np.pi * np.arcsin(                       # noqa: E731  kernels.py:145-151
                np.clip(2 * a / np.sqrt((1 + 2 * b) * (1 + 2 * c)), -1 + 1e-7, 1 - 1e-7))
        else:
            def f(a, b, c):                                                                   # kernels.py:178-183
                s = np.sqrt(b * c)
                fr = a / s
                th = np.arccos(np.clip(fr, -1 + 1e-7, 1 - 1e-7))
                return var_b + var_w / (2 * np.pi) * s * (np.sin(th) + (np.pi - th) * fr)
        k12, k11, k22 = f(k12, k11, k22), f(k11, k11, k11), f(k22, k22, k22)
    return k12


def nngp_kernel(X, Z, params, noise=0, jitter=1e-6, activation="erf", depth=3):
    d = X.shape[-1]
    xz = X @ Z.T
    xx = (X * X).sum(1)[:, None] * np.ones((1, Z.shape[0]))
    zz = np.ones((X.shape[0], 1)) * (Z * Z).sum(1)[None, :]
    k = _nngp_pair(xz, xx, zz, params["var_b"], params["var_w"], depth, activation, d)
    if X.shape == Z.shape:
        k = k + (noise + jitter) * np.eye(X.shape[0])                    # kernels.py:221-222
    return k


def measured_noise_logp(X, y, params, measured_noise, kernel="RBF", jitter=1e-6):
    k = go.get_kernel(kernel)(X, X, params, 0, jitter=jitter) + np.diag(measured_noise)      # mngp.py:92-97
    L = sla.cholesky(k, lower=True)
    a = sla.solve_triangular(L, y, lower=True)
    return -0.5 * a @ a - np.log(np.diag(L)).sum() - 0.5 * len(y) * np.log(2 * np.pi)


# ---- multi-task kernels (gpax/kernels/mtkernels.py)
def _task_cov(params):
    W, v = np.asarray(params["W"]), np.asarray(params["v"])
    return W @ W.T + np.diag(v)                                                              # mtkernels.py:55-57


def multitask_kernel(X, Z, params, noise, kernel="RBF", jitter=1e-6):
    """mtkernels.py:89-123: k_data(x, z) * B[task(x), task(z)], per-task noise + jitter on the diagonal when shapes agree."""
    Xd, tX, Zd, tZ = X[:, :-1], X[:, -1].astype(int), Z[:, :-1], Z[:, -1].astype(int)
    K = go.get_kernel(kernel)(Xd, Zd, params, 0, jitter=jitter)    # noise 0, but the data kernel's own diagonal rule adds `jitter` (:103)
    K = K * _task_cov(params)[np.ix_(tX, tZ)]
    if X.shape == Z.shape:
        K[np.diag_indices(len(X))] += np.asarray(noise)[tX] + jitter                         # :111-121
    return K


def multivariate_kernel(X, Z, params, noise, kernel="RBF", num_tasks=1, jitter=1e-6):
    """mtkernels.py:161-190: kron(k_data, k_task) + kron(I, diag(noise + jitter))."""
    K = np.kron(go.get_kernel(kernel)(X, Z, params, 0, jitter=jitter), _task_cov(params))
    if X.shape == Z.shape:
        K = K + np.kron(np.eye(len(X)), np.diag(np.asarray(noise) + jitter))
    return K


def lcm_kernel(X, Z, params, noise, kernel="RBF", jitter=1e-6):
    """mtkernels.py:226-230 (shared_input_space=False): sum of multi-task kernels over the leading axis of the parameters."""
    L = len(params["k_scale"])
    return sum(multitask_kernel(X, Z, {k: np.asarray(v)[q] for k, v in params.items()}, noise, kernel, jitter) for q in range(L))
