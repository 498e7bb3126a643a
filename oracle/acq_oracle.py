"""acq_oracle.py -- NumPy restatement of the reference's acquisition functions.  TEST INFRASTRUCTURE ONLY.

  ei / ucb / ue / poi        gpax/acquisition/base_acq.py:20-71, 74-104, 107-130, 133-155
  moments_from_samples       gpax/acquisition/acquisition.py:23-36 (_compute_mean_and_var, MCMC branch)
  kg                         gpax/acquisition/base_acq.py:158-232, LITERALLY: the training set is augmented with every
                             candidate and simulated observation and the posterior re-computed through the oracle's
                             explicit-inverse formulation (the CUDA path evaluates the block-inverse identity instead)

numpyro's Normal.cdf is jax.scipy.special.ndtr; exp(log_prob) is the density.  Pinned by tests/test_oracle_f.py against
tests/golden/reference_vectors_f.npz (the reference's own source run over the NumPy shim)."""
import numpy as np
from scipy.special import ndtr

from . import gp_oracle as go


def _pdf(u):
    return np.exp(-0.5 * u * u - 0.5 * np.log(2 * np.pi))


def ei(mean, var, best_f=None, maximize=False):
    if best_f is None:
        best_f = mean.max() if maximize else mean.min()          # base_acq.py:59-60
    sigma = np.sqrt(var)
    u = (mean - best_f) / sigma
    if not maximize:
        u = -u
    return sigma * (_pdf(u) + u * ndtr(u))                        # :66-69


def ucb(mean, var, beta=0.25, maximize=False):
    delta = np.sqrt(beta * var)
    return mean + delta if maximize else -(mean - delta)          # :98-103


def ue(mean, var):
    return np.sqrt(var)                                           # :129-130


def poi(mean, var, best_f=None, xi=0.01, maximize=False):
    if best_f is None:
        best_f = mean.max() if maximize else mean.min()
    u = (mean - best_f - xi) / np.sqrt(var)
    if not maximize:
        u = -u
    return ndtr(u)                                                # :150-155


def moments_from_samples(y_sampled):
    y = np.asarray(y_sampled)
    y = y.reshape(-1, y.shape[-1])
    return y.mean(0), y.var(0)                                    # acquisition.py:33-34


def kg(X_train, y_train, X_new, params, kernel, eps, maximize=True, noiseless=True, **kwargs):
    """base_acq.py:158-232 with the simulated observations y_sim = mean + eps @ chol(cov)^T (eps [n, P] injected)."""
    mean, cov = go.exact_posterior(X_train, y_train, X_new, params, kernel, noiseless, **kwargs)     # :219
    y_sim = mean[None, :] + eps @ np.linalg.cholesky(cov).T                                           # :221
    mean_o_best = mean.max() if maximize else mean.min()
    vals = np.empty((eps.shape[0], X_new.shape[0]))
    for i, ys in enumerate(y_sim):
        for c in range(X_new.shape[0]):
            Xa = np.concatenate([X_train, X_new[c][None]], axis=0)                                    # :223
            ya = np.concatenate([y_train, ys[c][None]])                                               # :226
            mean_aug, _ = go.exact_posterior(Xa, ya, X_new, params, kernel, noiseless, **kwargs)      # :206-207
            y_fant = mean_aug.max() if maximize else mean_aug.min()
            u = y_fant - mean_o_best
            vals[i, c] = u if maximize else -u
    return vals.mean(0)                                                                               # :232
