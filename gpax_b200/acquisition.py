"""
acquisition.py -- acquisition functions with the reference's surface: EI / UCB / POI / UE / KG
(gpax/acquisition/acquisition.py:50-500 over base_acq.py:20-232) and the q-batch forms qEI / qUCB / qPOI
(gpax/acquisition/batch_acquisition.py:60-260).  The arithmetic runs on the GPU as epilogues of the posterior
(b2gp_acq_moments / b2gp_acq_samples / b2gp_kg, gpax_b200/csrc/acq.cuh); the penalties of
gpax/acquisition/penalties.py are O(P * recent) host arithmetic and stay on the host, as in the reference.
"""
from typing import Optional

import numpy as np

from . import prng
from .utils import posterior_eps

__all__ = ["EI", "UCB", "POI", "UE", "KG", "qEI", "qUCB", "qPOI", "ei", "ucb", "poi", "ue", "compute_penalty"]


# ------------------------------------------------------------------ base functions on moments (base_acq.py:20-155)
def _ctx(model=None):
    from . import _ffi
    return model.ctx if model is not None else _ffi.default_context()


def ei(moments, best_f=None, maximize=False, ctx=None, **kwargs):
    mean, var = moments
    return (ctx or _ctx()).acq_moments("EI", mean, var, best_f, 0.0, maximize)


def ucb(moments, beta=0.25, maximize=False, ctx=None, **kwargs):
    mean, var = moments
    return (ctx or _ctx()).acq_moments("UCB", mean, var, None, beta, maximize)


def ue(moments, ctx=None, **kwargs):
    mean, var = moments
    return (ctx or _ctx()).acq_moments("UE", mean, var)


def poi(moments, best_f=None, xi=0.01, maximize=False, ctx=None, **kwargs):
    mean, var = moments
    return (ctx or _ctx()).acq_moments("POI", mean, var, best_f, xi, maximize)


# ------------------------------------------------------------------ penalties (penalties.py)
def _penalty_point(x, recent_points):
    if recent_points.ndim == 1:
        recent_points = recent_points[:, None]
    distances = np.linalg.norm(recent_points - x, axis=1)
    timestamps = 1 if len(recent_points) == 1 else np.arange(len(recent_points) + 1, 1, -1)
    return np.sum(1 / (distances + 1) / timestamps)


def compute_penalty(X, recent_points, penalty_type="delta", penalty_factor=1.0):
    """gpax/acquisition/penalties.py:15-45."""
    X, recent_points = np.asarray(X), np.asarray(recent_points)
    if penalty_type not in ["delta", "inverse_distance", "inverse distance"]:
        raise NotImplementedError("Avaialble penalty types are 'delta' and 'inverse distance'")
    if penalty_type == "delta":
        out = np.zeros(len(X))
        for single_point in recent_points:
            idx = np.where(np.all(X == single_point, axis=1))[0]
            if idx.size > 0:
                out[idx[0]] = np.inf
        return out
    return penalty_factor * np.array([_penalty_point(x, recent_points) for x in X])


def _penalise(acq, X, penalty, recent_points, grid_indices, penalty_factor):
    if penalty:
        X_ = grid_indices if grid_indices is not None else X
        acq = acq - compute_penalty(X_, recent_points, penalty, penalty_factor)
    return acq


def _check_penalty(penalty, recent_points):
    if penalty and not isinstance(recent_points, np.ndarray):
        raise ValueError("Please provide an array of recently visited points")


# ------------------------------------------------------------------ model-level functions (acquisition.py)
def _acq_for_model(kind, rng_key, model, X, n, noiseless, best_f, param, maximize, **kwargs):
    """acquisition.py:23-36 (_compute_mean_and_var) + the base function.  Fully Bayesian model: the moments are taken
    over the S*n posterior samples (column reduction on the device); viGP-style model: over (mean, var) directly."""
    if getattr(model, "mcmc", None) is not None:
        _, y_sampled = model.predict(rng_key, X, n=n, noiseless=noiseless, **kwargs)
        y = np.asarray(y_sampled, dtype=np.float64).reshape(n * y_sampled.shape[0], -1)
        acq, _, _ = model.ctx.acq_samples(kind, y, best_f, param, maximize)
        return acq
    mean, var = model.predict(rng_key, X, noiseless=noiseless, **kwargs)
    return model.ctx.acq_moments(kind, mean, var, best_f, param, maximize)


def EI(rng_key, model, X, best_f: float = None, maximize: bool = False, n: int = 1, noiseless: bool = False,
       penalty: Optional[str] = None, recent_points=None, grid_indices=None, penalty_factor: float = 1.0, **kwargs):
    """Expected improvement -- gpax/acquisition/acquisition.py:50-143."""
    _check_penalty(penalty, recent_points)
    X = np.asarray(X)
    X = X[:, None] if X.ndim < 2 else X
    acq = _acq_for_model("EI", rng_key, model, X, n, noiseless, best_f, 0.0, maximize, **kwargs)
    return _penalise(acq, X, penalty, recent_points, grid_indices, penalty_factor)


def UCB(rng_key, model, X, beta: float = 0.25, maximize: bool = False, n: int = 1, noiseless: bool = False,
        penalty: Optional[str] = None, recent_points=None, grid_indices=None, penalty_factor: float = 1.0, **kwargs):
    """Upper confidence bound -- acquisition.py:146-227."""
    _check_penalty(penalty, recent_points)
    X = np.asarray(X)
    X = X[:, None] if X.ndim < 2 else X
    acq = _acq_for_model("UCB", rng_key, model, X, n, noiseless, None, beta, maximize, **kwargs)
    return _penalise(acq, X, penalty, recent_points, grid_indices, penalty_factor)


def POI(rng_key, model, X, best_f: float = None, xi: float = 0.01, maximize: bool = False, n: int = 1,
        noiseless: bool = False, penalty: Optional[str] = None, recent_points=None, grid_indices=None,
        penalty_factor: float = 1.0, **kwargs):
    """Probability of improvement -- acquisition.py:230-314."""
    _check_penalty(penalty, recent_points)
    X = np.asarray(X)
    X = X[:, None] if X.ndim < 2 else X
    acq = _acq_for_model("POI", rng_key, model, X, n, noiseless, best_f, xi, maximize, **kwargs)
    return _penalise(acq, X, penalty, recent_points, grid_indices, penalty_factor)


def UE(rng_key, model, X, n: int = 1, noiseless: bool = False, penalty: Optional[str] = None, recent_points=None,
       grid_indices=None, penalty_factor: float = 1.0, **kwargs):
    """Uncertainty-based exploration -- acquisition.py:317-392."""
    _check_penalty(penalty, recent_points)
    X = np.asarray(X)
    X = X[:, None] if X.ndim < 2 else X
    acq = _acq_for_model("UE", rng_key, model, X, n, noiseless, None, 0.0, False, **kwargs)
    return _penalise(acq, X, penalty, recent_points, grid_indices, penalty_factor)


def kg(model, X_new, sample, rng_key=None, n: int = 10, maximize: bool = True, noiseless: bool = True, eps=None, **kwargs):
    """Knowledge gradient for one sample of the hyper-parameters -- base_acq.py:158-232.  One posterior call
    (mean, cov and the n simulated observations y_sim = mean + chol(cov) eps, base_acq.py:221-223) and one closed-form
    rank-1 update kernel instead of P*n re-factorisations.  `eps` [n, P] may be injected (tests); otherwise it follows
    the reference's key: normal(rng_key, (n, P))."""
    X_new = np.asarray(model._set_data(X_new), dtype=np.float64)
    P = X_new.shape[0]
    if eps is None:
        key = rng_key if rng_key is not None else prng.PRNGKey(0)
        eps = posterior_eps(key, 1, n, P, np.float32, per_draw_keys=False)
    out = model._posterior_batched(X_new, sample, False, noiseless, ("mean", "cov"), eps=np.asarray(eps).reshape(1, n, P), **kwargs)
    mean, cov, ysim = out["mean"][0], out["cov"][0], out["y_sampled"][0]
    noise = float(np.asarray(sample["noise"]))
    jitter = float(kwargs.get("jitter", 1e-6))
    diag_sub = noise * (0.0 if noiseless else 1.0) + jitter
    return model.ctx.kg(mean, cov, ysim, diag_sub, noise + jitter, maximize)


def KG(rng_key, model, X, n: int = 1, maximize: bool = False, noiseless: bool = False, penalty: Optional[str] = None,
       recent_points=None, grid_indices=None, penalty_factor: float = 1.0, **kwargs):
    """Knowledge gradient -- acquisition.py:395-484: kg() for the variational model's parameters, or one row per
    posterior draw ([S, P], the reference's vmap over the draws) for an MCMC model."""
    _check_penalty(penalty, recent_points)
    X = np.asarray(X)
    X = X[:, None] if X.ndim < 2 else X
    samples = model.get_samples()
    if getattr(model, "mcmc", None) is not None:
        S = len(next(iter(samples.values())))
        keys = prng.split(prng.as_key(rng_key), S)
        vals = []
        for s in range(S):
            one = {k: np.asarray(v)[s] for k, v in samples.items()}
            vals.append(kg(model, X, one, keys[s], n, maximize, noiseless, **kwargs))
        acq = np.stack(vals)
    else:
        acq = kg(model, X, samples, rng_key, n, maximize, noiseless, **kwargs)
    return _penalise(acq, X, penalty, recent_points, grid_indices, penalty_factor)


# ------------------------------------------------------------------ q-batch functions (batch_acquisition.py)
def _subsample(samples, num, rng_key):
    """gpax/utils/utils.py:84-102 (random_sample_dict): `num` consistent rows of every site."""
    N = len(next(iter(samples.values())))
    rng = np.random.default_rng(int(np.asarray(prng.as_key(rng_key), dtype=np.uint64).sum()))
    idx = rng.permutation(N)[:num]
    return {k: np.asarray(v)[idx] for k, v in samples.items()}


def _q_acq(kind, rng_key, model, X, best_f, param, maximize, noiseless, maximize_distance, subsample_size, n_evals,
           indices, **kwargs):
    """batch_acquisition.py:20-57: the acquisition function of `subsample_size` individual posterior draws, one row each
    ([subsample_size, P]); one batched posterior call (mean + diag variance) and one epilogue launch per evaluation."""
    if getattr(model, "mcmc", None) is None:
        raise ValueError("The model needs to be fully Bayesian")
    X = np.asarray(X)
    X = X[:, None] if X.ndim < 2 else X

    def rows(samples, Xq):
        out = model._posterior_batched(Xq, samples, True, noiseless, ("mean", "var"), **kwargs)
        return model.ctx.acq_moments(kind, out["mean"], out["var"], best_f, param, maximize)

    if not maximize_distance:
        return rows(_subsample(model.get_samples(), subsample_size, rng_key), X)
    X_ = np.asarray(indices) if indices is not None else X
    best, best_d = None, -np.inf
    for sub in prng.split(prng.as_key(rng_key), n_evals):
        acq = rows(_subsample(model.get_samples(), subsample_size, sub), X_)
        d = np.linalg.norm(acq.argmax(-1)).mean()
        if d > best_d:
            best, best_d = acq, d
    return best


def qEI(rng_key, model, X, best_f: float = None, maximize: bool = False, noiseless: bool = False,
        maximize_distance: bool = False, subsample_size: int = 1, n_evals: int = 10, indices=None, **kwargs):
    """batch_acquisition.py:60-118."""
    return _q_acq("EI", rng_key, model, X, best_f, 0.0, maximize, noiseless, maximize_distance, subsample_size, n_evals,
                  indices, **kwargs)


def qUCB(rng_key, model, X, beta: float = 0.25, maximize: bool = False, noiseless: bool = False,
         maximize_distance: bool = False, subsample_size: int = 1, n_evals: int = 10, indices=None, **kwargs):
    """batch_acquisition.py:121-175."""
    return _q_acq("UCB", rng_key, model, X, None, beta, maximize, noiseless, maximize_distance, subsample_size, n_evals,
                  indices, **kwargs)


def qPOI(rng_key, model, X, best_f: float = None, xi: float = 0.01, maximize: bool = False, noiseless: bool = False,
         maximize_distance: bool = False, subsample_size: int = 1, n_evals: int = 10, indices=None, **kwargs):
    """batch_acquisition.py:178-232."""
    return _q_acq("POI", rng_key, model, X, best_f, xi, maximize, noiseless, maximize_distance, subsample_size, n_evals,
                  indices, **kwargs)
