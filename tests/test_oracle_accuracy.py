"""How far the fp64 oracles themselves are from the mathematical posterior (60-digit arithmetic, oracle/mp_oracle.py):
the number that decides what a disagreement of 1e-9 between the CUDA path and the oracle can mean."""
import numpy as np
import pytest

import oracle
from oracle import mp_oracle


def _case(kname, noise, seed):
    rng = np.random.default_rng(seed)
    N, P, d = 40, 5, 2
    X = rng.uniform(0, 1, (N, d))
    y = np.sin(4 * X[:, 0]) + X[:, 1] + 0.05 * rng.standard_normal(N)
    Xn = rng.uniform(0, 1, (P, d))
    params = {"k_length": np.array([0.4, 0.6]), "k_scale": 1.3, "noise": noise, "period": 0.8}
    K = oracle.get_kernel(kname)(X, X, params, noise)
    return X, y, Xn, params, np.linalg.cond(K)


@pytest.mark.parametrize("kname", ["RBF", "Matern", "Periodic"])
@pytest.mark.parametrize("noise", [0.1, 1e-3])
def test_fp64_oracles_against_60_digit_posterior(kname, noise):
    X, y, Xn, params, cond = _case(kname, noise, 3)
    mean_mp, cov_mp = mp_oracle.exact_posterior_mp(X, y, Xn, params, kname)
    eps = np.finfo(np.float64).eps
    for fn in (oracle.exact_posterior, oracle.exact_posterior_chol):
        mean, cov = fn(X, y, Xn, params, kname)
        em = np.abs(mean - mean_mp).max() / np.abs(mean_mp).max()
        ec = np.abs(cov - cov_mp).max() / np.abs(cov_mp).max()
        # both formulations are backward stable up to the conditioning of K: a few hundred cond * eps at most
        bound = 500 * cond * eps
        assert em <= bound and ec <= bound, (fn.__name__, kname, noise, cond, em, ec)
    # at the benchmark's conditioning (cond <= 1e5) the oracle is good to ~1e-9, the parity tolerance of the GPU tests
    if cond <= 1e5:
        mean, cov = oracle.exact_posterior(X, y, Xn, params, kname)
        assert np.abs(mean - mean_mp).max() <= 1e-9 * np.abs(mean_mp).max()
        assert np.abs(cov - cov_mp).max() <= 1e-9 * np.abs(cov_mp).max()
