"""Prior programs in fit() (gpax/models/gp.py:137-154; the programs of tests/test_gp.py:25-38 with one changed import).
CPU: the likelihood is a NumPy restatement handed to ProgramLogJoint in place of the GPU call, so the host logic -- site
discovery, transforms, the chain rule through the programs, hierarchical priors -- is tested without a device."""
import numpy as np
import pytest

from gpax_b200 import ExactGP, priors as numpyro
from gpax_b200.inference import LogJoint, ProgramLogJoint, make_log_joint, run_nuts


def dummy_mean_fn(x, params):                     # tests/test_gp.py:25-26
    return params["a"] * x ** params["b"]


def dummy_mean_fn_priors():                       # tests/test_gp.py:29-32
    a = numpyro.sample("a", numpyro.distributions.LogNormal(0, 1))
    b = numpyro.sample("b", numpyro.distributions.Normal(3, 1))
    return {"a": a, "b": b}


def gp_kernel_custom_prior():                     # tests/test_gp.py:35-38
    length = numpyro.sample("k_length", numpyro.distributions.Uniform(0, 1))
    scale = numpyro.sample("k_scale", numpyro.distributions.LogNormal(0, 1))
    return {"k_length": length, "k_scale": scale}


def hierarchical_prior():
    top = numpyro.sample("top", numpyro.distributions.LogNormal(0, 0.5))
    with numpyro.plate("ard", 2):
        length = numpyro.sample("k_length", numpyro.distributions.LogNormal(np.log(top), 0.3))
    return {"k_length": length, "k_scale": numpyro.deterministic("k_scale", 2.0 * top)}


def numpy_lik(X):
    """exact RBF marginal likelihood, d/dlog(theta), alpha -- what b2gp_mll returns"""
    N, d = X.shape

    def lik(th, yres):
        ell, scale, noise = th[:d], th[d], th[d + 1]
        D = ((X[:, None, :] - X[None, :, :]) / ell) ** 2
        K0 = scale * np.exp(-0.5 * D.sum(-1))
        K = K0 + (noise + 1e-6) * np.eye(N)
        L = np.linalg.cholesky(K)
        alpha = np.linalg.solve(K, yres)
        val = -0.5 * yres @ alpha - np.log(np.diag(L)).sum() - 0.5 * N * np.log(2 * np.pi)
        W = np.outer(alpha, alpha) - np.linalg.inv(K)
        g = np.zeros(d + 3)
        for i in range(d):
            g[i] = 0.5 * np.sum(W * K0 * D[:, :, i])
        g[d] = 0.5 * np.sum(W * K0)
        g[d + 1] = 0.5 * noise * np.trace(W)
        return val, g, alpha, 0
    return lik


def data(d=1, N=12, seed=0):
    rng = np.random.default_rng(seed)
    X = np.sort(rng.uniform(1, 2, (N, d)), axis=0)
    y = 10 * X[:, 0] ** 2 + 0.1 * rng.standard_normal(N)
    return X, y


def fd_grad(lj, u, jac, h=1e-6):
    g = np.zeros_like(u)
    for k in range(u.size):
        e = np.zeros_like(u)
        e[k] = h
        g[k] = (lj(u + e, jac)[0] - lj(u - e, jac)[0]) / (2 * h)
    return g


@pytest.mark.parametrize("jac", [False, True])
def test_reference_programs_gradient(jac):
    X, y = data()
    m = ExactGP(1, "RBF", mean_fn=dummy_mean_fn, mean_fn_prior=dummy_mean_fn_priors, kernel_prior=gp_kernel_custom_prior)
    m.X_train, m.y_train = X, y
    lj = ProgramLogJoint(m, lik=numpy_lik(X))
    assert [s.name for s in lj.sites] == ["k_length", "k_scale", "noise", "a", "b"] and lj.dim == 5
    assert not lj.hierarchical
    u = np.array([0.3, -0.2, -1.0, 2.0, 2.1])
    val, g = lj(u, jac)
    assert np.isfinite(val)
    np.testing.assert_allclose(g, fd_grad(lj, u, jac), rtol=2e-6, atol=1e-6)
    # the value: likelihood at the transformed sites + site densities (+ log-Jacobians)
    ell = 1 / (1 + np.exp(-0.3))
    th = np.array([ell, np.exp(-0.2), np.exp(-1.0), 1.0])
    want = numpy_lik(X)(th, y - np.exp(2.0) * X[:, 0] ** 2.1)[0]
    want += -np.log(1.0) + float(numpyro.LogNormal(0, 1).log_prob(th[1])) + float(numpyro.LogNormal(0, 1).log_prob(th[2]))
    want += float(numpyro.LogNormal(0, 1).log_prob(np.exp(2.0))) + float(numpyro.Normal(3, 1).log_prob(2.1))
    if jac:
        want += np.log(ell * (1 - ell)) - 0.2 - 1.0 + 2.0
    assert abs(val - want) < 1e-9 * max(1.0, abs(want))


def test_hierarchical_program_gradient_and_shapes():
    X, y = data(d=2)
    m = ExactGP(2, "RBF", kernel_prior=hierarchical_prior)
    m.X_train, m.y_train = X, y
    lj = make_log_joint(m)
    assert isinstance(lj, ProgramLogJoint)
    lj._lik_fn = numpy_lik(X)
    assert lj.hierarchical and lj.dim == 4            # top, k_length[2], noise
    u = np.array([0.2, -0.5, 0.1, -1.5])
    np.testing.assert_allclose(lj(u, True)[1], fd_grad(lj, u, True), rtol=2e-6, atol=1e-6)
    out = lj.to_dict(np.stack([u, u + 0.1, u - 0.1]))
    assert out["k_length"].shape == (3, 2) and out["top"].shape == (3,) and out["noise"].shape == (3,)


def test_default_program_equals_the_analytic_log_joint():
    X, y = data(d=2)
    m = ExactGP(2, "RBF", lengthscale_prior_dist=numpyro.Gamma(2, 5), noise_prior_dist=numpyro.HalfNormal(0.1))
    m.X_train, m.y_train = X, y

    class CpuLJ(LogJoint):
        def _lik(self, th):
            v, g, _, info = numpy_lik(X)(th, self.y)
            return v, g, info
    a = CpuLJ(m)
    b = ProgramLogJoint(m, lik=numpy_lik(X))
    u = np.array([-0.4, 0.2, 0.3, -2.0])
    for jac in (False, True):
        va, ga = a(u, jac)
        vb, gb = b(u, jac)
        assert abs(va - vb) < 1e-10 * abs(va)
        np.testing.assert_allclose(gb, ga, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(a.init_u(), b.init_u())


def test_nuts_over_a_program_log_joint_recovers_the_mean_function():
    X, y = data(N=16)
    m = ExactGP(1, "RBF", mean_fn=dummy_mean_fn, mean_fn_prior=dummy_mean_fn_priors)
    m.X_train, m.y_train = X, y
    lj = ProgramLogJoint(m, lik=numpy_lik(X))
    res = run_nuts(lj, 0, 150, 100, 1, False)
    s = res.get_samples()
    assert set(s) == {"k_length", "k_scale", "noise", "a", "b"} and s["k_length"].shape == (100, 1) and s["a"].shape == (100,)
    fit = np.mean([dummy_mean_fn(X[:, 0], {"a": a, "b": b}) for a, b in zip(s["a"], s["b"])], axis=0)
    assert np.sqrt(np.mean((fit - y) ** 2)) < 0.25 * np.std(y)      # y = 10 x^2: the parametric mean carries the trend


def test_program_errors():
    with pytest.raises(RuntimeError):
        numpyro.sample("a", numpyro.LogNormal(0, 1))                  # outside a program

    def twice():
        numpyro.sample("a", numpyro.LogNormal(0, 1))
        return numpyro.sample("a", numpyro.LogNormal(0, 1))
    with pytest.raises(ValueError):
        numpyro.run_program(twice)
    with pytest.raises(TypeError):
        numpyro.run_program(lambda: numpyro.sample("a", object()))


def test_prior_draws_follow_the_priors():
    """sample_from_prior's host half (gp.py:401-408): every site drawn from its prior, programs included"""
    from gpax_b200.inference import prior_draws
    m = ExactGP(2, "Periodic", lengthscale_prior_dist=numpyro.Gamma(2, 5), noise_prior_dist=numpyro.HalfNormal(0.1))
    dr = prior_draws(m, np.random.default_rng(1), 4000, 2)
    ell = np.array([kp["k_length"] for kp, _, _ in dr])
    noise = np.array([n for _, n, _ in dr])
    period = np.array([kp["period"] for kp, _, _ in dr])
    assert ell.shape == (4000, 2) and abs(ell.mean() - 0.4) < 0.02            # Gamma(2, 5): mean 0.4
    assert abs(noise.mean() - 0.1 * np.sqrt(2 / np.pi)) < 0.005               # HalfNormal(0.1)
    assert abs(np.median(period) - 1.0) < 0.08                                 # LogNormal(0, 1): median 1
    m2 = ExactGP(1, "RBF", mean_fn=dummy_mean_fn, mean_fn_prior=dummy_mean_fn_priors, kernel_prior=hierarchical_prior_1d)
    kp, noise, mp = prior_draws(m2, np.random.default_rng(2), 1, 1)[0]
    assert set(mp) == {"a", "b"} and kp["k_scale"] > 0 and np.shape(kp["k_length"]) == (1,)


def hierarchical_prior_1d():
    top = numpyro.sample("top", numpyro.distributions.LogNormal(0, 0.5))
    with numpyro.plate("ard", 1):
        length = numpyro.sample("k_length", numpyro.distributions.LogNormal(np.log(top), 0.3))
    return {"k_length": length, "k_scale": 2.0 * top}
