"""CPU: the host-side samplers / optimiser of gpax_b200/inference.py on analytic targets (no GPU, no oracle)."""
import math

import numpy as np

from gpax_b200 import inference as inf
from gpax_b200 import priors as P


class GaussTarget:
    """log N(u; mu, diag(sd^2)) with the LogJoint call signature"""
    def __init__(self, mu, sd):
        self.mu, self.sd, self.dim, self.n_evals = np.asarray(mu, float), np.asarray(sd, float), len(mu), 0

    def __call__(self, u, jacobian):
        self.n_evals += 1
        z = (u - self.mu) / self.sd
        return float(-0.5 * np.sum(z * z)), -z / self.sd


def test_nuts_samples_a_gaussian():
    tgt = GaussTarget([1.0, -2.0, 0.5], [0.5, 2.0, 1.0])
    rng = np.random.default_rng(0)
    u = np.zeros(3)
    lp, g = tgt(u, True)
    minv = np.ones(3)
    eps = inf._find_eps(tgt, u, lp, g, rng, minv)
    assert 0.05 < eps < 8.0
    draws = []
    for it in range(1500):
        u, lp, g, acc, depth, div = inf._nuts_draw(tgt, u, lp, g, 0.6, rng, minv)
        assert not div and 0.0 <= acc <= 1.0
        if it >= 300:
            draws.append(u.copy())
    d = np.array(draws)
    np.testing.assert_allclose(d.mean(0), tgt.mu, atol=0.25)
    np.testing.assert_allclose(d.std(0), tgt.sd, rtol=0.2)


def test_prior_transforms_roundtrip_and_gradients():
    for pr in (P.LogNormal(0.3, 0.7), P.HalfNormal(0.5), P.Gamma(2.0, 5.0), P.Uniform(0.1, 3.0)):
        u, h = np.array(0.37), 1e-6
        assert abs(float(pr.inverse(pr.transform(u))) - 0.37) < 1e-12
        num = (pr.log_prob(pr.transform(u + h)) - pr.log_prob(pr.transform(u - h))) / (2 * h)
        assert abs(float(num) - float(pr.dlog_prob(pr.transform(u)) * pr.dtheta_du(u))) < 1e-6
        numj = (pr.log_abs_jac(u + h) - pr.log_abs_jac(u - h)) / (2 * h)
        assert abs(float(numj) - float(pr.dlog_abs_jac(u))) < 1e-6
        med = pr.median()
        s = pr.sample(np.random.default_rng(1), (20001,))
        assert abs(np.median(s) - med) < 0.05 * max(1.0, abs(med))


def test_lognormal_default_prior_is_standard_normal_in_u():
    pr = P.LogNormal(0.0, 1.0)
    u = np.linspace(-2, 2, 9)
    lp_u = pr.log_prob(pr.transform(u)) + pr.log_abs_jac(u)
    np.testing.assert_allclose(lp_u, -0.5 * u * u - 0.5 * math.log(2 * math.pi), rtol=1e-13)
