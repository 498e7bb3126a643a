"""
priors.py -- minimal prior distributions for `.fit()` (the reference passes numpyro.distributions objects:
gpax/models/gp.py:222-247 defaults to LogNormal(0, 1) on k_length, k_scale, noise, period).  Each prior carries
its support transform theta(u) so that inference runs on an unconstrained vector u, as NumPyro does.
"""
import math

import numpy as np


class Prior:
    """theta = transform(u); log_prob is the density over theta."""
    positive = True

    def transform(self, u):            # theta(u)
        return np.exp(u)

    def dtheta_du(self, u):
        return np.exp(u)

    def log_abs_jac(self, u):          # log |dtheta/du|
        return u

    def dlog_abs_jac(self, u):
        return np.ones_like(u)

    def inverse(self, theta):
        return np.log(theta)

    def median(self):
        raise NotImplementedError

    def log_prob(self, theta):
        raise NotImplementedError

    def dlog_prob(self, theta):
        raise NotImplementedError

    def sample(self, rng, shape=()):
        raise NotImplementedError


class LogNormal(Prior):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = float(loc), float(scale)

    def median(self):
        return math.exp(self.loc)

    def log_prob(self, t):
        z = (np.log(t) - self.loc) / self.scale
        return -np.log(t) - 0.5 * z * z - math.log(self.scale) - 0.5 * math.log(2 * math.pi)

    def dlog_prob(self, t):
        return (-1.0 - (np.log(t) - self.loc) / self.scale ** 2) / t

    def sample(self, rng, shape=()):
        return np.exp(self.loc + self.scale * rng.standard_normal(shape))


class HalfNormal(Prior):
    def __init__(self, scale=1.0):
        self.scale = float(scale)

    def median(self):
        return 0.6744897501960817 * self.scale

    def log_prob(self, t):
        return 0.5 * math.log(2 / math.pi) - math.log(self.scale) - 0.5 * (t / self.scale) ** 2

    def dlog_prob(self, t):
        return -t / self.scale ** 2

    def sample(self, rng, shape=()):
        return np.abs(self.scale * rng.standard_normal(shape))


class Gamma(Prior):
    def __init__(self, concentration, rate=1.0):
        self.a, self.b = float(concentration), float(rate)

    def median(self):
        from scipy.stats import gamma
        return float(gamma.median(self.a, scale=1.0 / self.b))

    def log_prob(self, t):
        return self.a * math.log(self.b) - math.lgamma(self.a) + (self.a - 1) * np.log(t) - self.b * t

    def dlog_prob(self, t):
        return (self.a - 1) / t - self.b

    def sample(self, rng, shape=()):
        return rng.gamma(self.a, 1.0 / self.b, shape)


class Uniform(Prior):
    def __init__(self, low, high):
        self.low, self.high = float(low), float(high)

    def _sig(self, u):
        return 1.0 / (1.0 + np.exp(-u))

    def transform(self, u):
        return self.low + (self.high - self.low) * self._sig(u)

    def dtheta_du(self, u):
        s = self._sig(u)
        return (self.high - self.low) * s * (1 - s)

    def log_abs_jac(self, u):
        s = self._sig(u)
        return math.log(self.high - self.low) + np.log(s) + np.log1p(-s)

    def dlog_abs_jac(self, u):
        return 1.0 - 2.0 * self._sig(u)

    def inverse(self, theta):
        p = (theta - self.low) / (self.high - self.low)
        return np.log(p) - np.log1p(-p)

    def median(self):
        return 0.5 * (self.low + self.high)

    def log_prob(self, t):
        return np.full_like(np.asarray(t, dtype=float), -math.log(self.high - self.low))

    def dlog_prob(self, t):
        return np.zeros_like(np.asarray(t, dtype=float))

    def sample(self, rng, shape=()):
        return rng.uniform(self.low, self.high, shape)
