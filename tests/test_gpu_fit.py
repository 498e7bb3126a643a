"""GPU: the fit path: b2gp_mll value / gradient, viGP.fit (SVI, Adam b1=0.5), ExactGP.fit (NUTS)."""
import numpy as np
import pytest
import scipy.linalg as sla

import oracle

pytestmark = pytest.mark.gpu

KMAP = {"RBF": oracle.rbf_kernel, "Matern": oracle.matern_kernel, "Periodic": oracle.periodic_kernel}


def mll_ref(kind, X, y, theta, jitter):
    d = X.shape[1]
    params = {"k_length": theta[:d], "k_scale": theta[d], "period": theta[d + 2]}
    K = KMAP[kind](X, X, params, theta[d + 1], jitter=jitter)
    L = sla.cholesky(K, lower=True)
    w = sla.solve_triangular(L, y, lower=True)
    return -0.5 * w @ w - np.log(np.diag(L)).sum() - 0.5 * len(y) * np.log(2 * np.pi)


@pytest.mark.parametrize("kind,d,N", [("RBF", 1, 50), ("Matern", 2, 300), ("Periodic", 1, 130), ("RBF", 3, 700), ("Matern", 2, 2300)])
def test_mll_value_and_gradient(kind, d, N):
    import gpax_b200
    ctx = gpax_b200.default_context()
    rng = np.random.default_rng(N)
    X = rng.uniform(0, 2, (N, d))
    y = np.sin(3 * X[:, 0]) + 0.2 * rng.standard_normal(N)
    theta = np.concatenate([rng.uniform(0.4, 0.9, d), [1.3, 0.15, 1.1]])
    val, grad, alpha, info = ctx.mll(kind, X, y, theta, 1e-6, want_grad=True, want_alpha=True)
    assert info == 0
    ref = mll_ref(kind, X, y, theta, 1e-6)
    assert abs(val - ref) <= 1e-9 * abs(ref)
    # alpha = K^-1 y
    params = {"k_length": theta[:d], "k_scale": theta[d], "period": theta[d + 2]}
    K = KMAP[kind](X, X, params, theta[d + 1], jitter=1e-6)
    np.testing.assert_allclose(alpha, np.linalg.solve(K, y), rtol=1e-7, atol=1e-8 * np.abs(alpha).max())
    # gradient w.r.t. log theta by central differences of the reference value
    idx = list(range(d + 2)) + ([d + 2] if kind == "Periodic" else [])
    for k in idx:
        h = 1e-5
        tp, tm = theta.copy(), theta.copy()
        tp[k] *= np.exp(h)
        tm[k] *= np.exp(-h)
        num = (mll_ref(kind, X, y, tp, 1e-6) - mll_ref(kind, X, y, tm, 1e-6)) / (2 * h)
        assert abs(grad[k] - num) <= 2e-5 * max(1.0, abs(num)), (k, grad[k], num)
    if kind != "Periodic":
        assert grad[d + 2] == 0.0


def test_mll_not_positive_definite():
    import gpax_b200
    X = np.linspace(0, 1, 40)[:, None]
    val, grad, _, info = gpax_b200.default_context().mll("RBF", X, np.ones(40), np.array([0.3, -1.0, 0.1, 1.0]))
    assert info > 0 and np.isnan(val) and np.isnan(grad).all()


def make_data(n=60, seed=0):
    rng = np.random.default_rng(seed)
    X = np.sort(rng.uniform(0, 4, n))
    y = np.sin(2 * X) + 0.1 * rng.standard_normal(n)
    return X, y


@pytest.mark.parametrize("guide", ["delta", "normal"])
def test_vigp_fit_then_predict(guide):
    """tests/test_vigp.py:27-65 contract (fit sets svi / kernel_params, predict works) + the objective decreases"""
    import gpax_b200
    X, y = make_data()
    m = gpax_b200.viGP(1, "RBF", guide=guide)
    m.fit(0, X, y, num_steps=300, step_size=0.05, progress_bar=False, print_summary=False)
    assert m.svi is not None and set(m.kernel_params) == {"k_length", "k_scale", "noise"}
    assert m.kernel_params["k_length"].shape == (1,)
    assert m.svi.losses[-20:].mean() < m.svi.losses[:20].mean()
    p = m.get_samples()
    assert 0.001 < p["noise"] < 0.1 and 0.3 < p["k_length"][0] < 3.0          # true noise var 0.01, sin(2x)
    Xt = np.linspace(0, 4, 50)
    mean, var = m.predict(0, Xt, noiseless=True)
    assert np.abs(mean - np.sin(2 * Xt)).max() < 0.2 and (var > 0).all()
    # deterministic given the key
    m2 = gpax_b200.viGP(1, "RBF", guide=guide)
    m2.fit(0, X, y, num_steps=300, step_size=0.05, progress_bar=False, print_summary=False)
    np.testing.assert_array_equal(m2.kernel_params["k_scale"], m.kernel_params["k_scale"])


def test_vigp_map_is_a_stationary_point():
    """the delta-guide optimum: gradient of log p(y|theta) + log LogNormal(theta) vanishes (vigp.py:108-120 semantics)"""
    import gpax_b200
    from gpax_b200.inference import LogJoint
    X, y = make_data(40, 3)
    m = gpax_b200.viGP(1, "Matern")
    m.fit(1, X, y, num_steps=1500, step_size=0.02, progress_bar=False, print_summary=False)
    lj = LogJoint(m)
    val, g = lj(m.svi.loc, jacobian=False)
    assert np.abs(g).max() < 5e-2


def test_exactgp_fit_nuts():
    """tests/test_gp.py:52-77 contract: fit populates mcmc; sample dict shapes; same key -> same samples"""
    import gpax_b200
    X, y = make_data(30, 1)
    m = gpax_b200.ExactGP(1, "RBF")
    m.fit(0, X, y, num_warmup=60, num_samples=40, progress_bar=False, print_summary=False)
    s = m.get_samples()
    assert s["k_length"].shape == (40, 1) and s["k_scale"].shape == (40,) and s["noise"].shape == (40,)
    assert all((v > 0).all() for v in s.values())
    s1 = m.get_samples(chain_dim=True)
    assert s1["k_length"].shape == (1, 40, 1)
    m2 = gpax_b200.ExactGP(1, "RBF")
    m2.fit(0, X, y, num_warmup=60, num_samples=40, progress_bar=False, print_summary=False)
    np.testing.assert_array_equal(m2.get_samples()["noise"], s["noise"])
    Xt = np.linspace(0, 4, 25)
    ymean, ysamp = m.predict(1, Xt, n=2)
    assert ymean.shape == (25,) and ysamp.shape == (40, 2, 25) and np.isfinite(ysamp).all()
    assert np.abs(ymean - np.sin(2 * Xt)).max() < 0.35
    # posterior noise should concentrate well below the LogNormal(0,1) prior median of 1
    assert np.median(s["noise"]) < 0.2


def test_predict_in_batches_reuses_factor():
    """the reference re-inverts k_XX per chunk (gp.py:319-322); here chunks 2.. hit the factor cache"""
    import ctypes
    import gpax_b200
    X, y = make_data(200, 5)
    params = {"k_length": np.array([0.7]), "k_scale": 1.0, "noise": 0.05}
    m = gpax_b200.viGP(1, "Matern")
    m.X_train, m.y_train = X, y
    Xt = np.linspace(0, 4, 103)
    fn = m.ctx.lib.b2gp_debug_cache_hits
    fn.restype, fn.argtypes = ctypes.c_int64, [ctypes.c_void_p]
    m.ctx.set_option("drop_factor_cache", 1)
    h0 = fn(m.ctx.h)
    mean, var = m.predict(0, Xt, params)
    m.INTERNAL_BATCH = 1                          # the caller's chunks as they are: 11 library calls, 11 cache hits
    mb, vb = m.predict_in_batches(0, Xt, 10, params)
    assert fn(m.ctx.h) - h0 >= 10
    np.testing.assert_allclose(mb, mean, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(vb, var, rtol=1e-11, atol=1e-13)
    del m.INTERNAL_BATCH                          # default: chunks below 8192 rows are merged -> one call, same outputs
    h1 = fn(m.ctx.h)
    mc, vc = m.predict_in_batches(0, Xt, 10, params)
    assert fn(m.ctx.h) - h1 == 1
    np.testing.assert_array_equal(mc, mean)
    np.testing.assert_array_equal(vc, var)
    # a different training set must not reuse the factor
    m.X_train = X + 0.01
    mean2, _ = m.predict(0, Xt, params)
    ref, _ = oracle.vi_predict(X + 0.01, y, Xt, params, "Matern")
    np.testing.assert_allclose(mean2, ref, rtol=1e-8, atol=1e-9)


# ------------------------------------------------------------------ sparse GP: VFE bound and its gradients
def elbo_ref(kind, X, y, Xu, theta, jitter):
    """NumPy restatement of viSparseGP.model's score (gpax/models/sparse_gp.py:91-114) for a small problem"""
    from scipy.stats import multivariate_normal
    d = X.shape[1]
    params = {"k_length": theta[:d], "k_scale": theta[d], "period": theta[d + 2]}
    noise = theta[d + 1]
    k = KMAP[kind]
    Luu = sla.cholesky(k(Xu, Xu, params, 0.0, jitter=jitter), lower=True)
    W = sla.solve_triangular(Luu, k(Xu, X, params, 0.0, jitter=0.0), lower=True)
    kd = k(X[:1], X[:1], params, 0.0, jitter=0.0)[0, 0]
    trace_term = max((len(y) * kd - (W ** 2).sum()) / noise, 0.0)
    S = W.T @ W + noise * np.eye(len(y))
    return multivariate_normal(np.zeros(len(y)), S).logpdf(y) - 0.5 * trace_term


@pytest.mark.parametrize("kind,d,N,M", [("RBF", 1, 60, 7), ("Matern", 2, 150, 20), ("Periodic", 1, 80, 9)])
def test_sparse_elbo_value_and_gradients(kind, d, N, M):
    import gpax_b200
    ctx = gpax_b200.default_context()
    rng = np.random.default_rng(N + M)
    X = rng.uniform(0, 2, (N, d))
    y = np.sin(3 * X[:, 0]) + 0.2 * rng.standard_normal(N)
    Xu = X[rng.choice(N, M, replace=False)] + 0.01 * rng.standard_normal((M, d))
    theta = np.concatenate([rng.uniform(0.5, 0.9, d), [1.3, 0.15, 1.3]])
    val, g, gx, info = ctx.sparse_elbo(kind, Xu, X, y, theta, 1e-5)
    assert info == 0
    ref = elbo_ref(kind, X, y, Xu, theta, 1e-5)
    assert abs(val - ref) <= 1e-8 * abs(ref), (val, ref)
    idx = list(range(d + 2)) + ([d + 2] if kind == "Periodic" else [])
    h = 1e-5
    for k in idx:
        tp, tm = theta.copy(), theta.copy()
        tp[k] *= np.exp(h)
        tm[k] *= np.exp(-h)
        num = (elbo_ref(kind, X, y, Xu, tp, 1e-5) - elbo_ref(kind, X, y, Xu, tm, 1e-5)) / (2 * h)
        assert abs(g[k] - num) <= 1e-4 * max(1.0, abs(num)), ("theta", k, g[k], num)
    for (a, k) in [(0, 0), (M // 2, d - 1), (M - 1, 0)]:
        Xp, Xm = Xu.copy(), Xu.copy()
        Xp[a, k] += h
        Xm[a, k] -= h
        num = (elbo_ref(kind, X, y, Xp, theta, 1e-5) - elbo_ref(kind, X, y, Xm, theta, 1e-5)) / (2 * h)
        assert abs(gx[a, k] - num) <= 1e-4 * max(1.0, abs(num)), ("Xu", a, k, gx[a, k], num)


def test_visparsegp_fit_then_predict():
    """tests/test_sparsegp.py:26-44 contract: fit sets m.Xu, more steps move it; prediction is sensible"""
    import gpax_b200
    rng = np.random.default_rng(0)
    X = np.sort(rng.uniform(0, 4, 200))
    y = np.sin(2 * X) + 0.1 * rng.standard_normal(200)
    m = gpax_b200.viSparseGP(1, "RBF")
    m.fit(0, X, y, inducing_points_ratio=0.1, num_steps=1, step_size=0.02, progress_bar=False, print_summary=False)
    Xu1 = np.array(m.Xu)
    assert Xu1.shape == (20, 1)
    m.fit(0, X, y, inducing_points_ratio=0.1, num_steps=300, step_size=0.02, progress_bar=False, print_summary=False)
    assert not np.allclose(m.Xu, Xu1)
    assert m.svi.losses[-10:].mean() < m.svi.losses[:10].mean()
    assert set(m.kernel_params) == {"k_length", "k_scale", "noise"}
    Xt = np.linspace(0.2, 3.8, 40)
    mean, var = m.predict(0, Xt, noiseless=True)
    assert np.abs(mean - np.sin(2 * Xt)).max() < 0.25 and (var > 0).all()
    mean2, cov2 = m.get_mvn_posterior(Xt, m.get_samples(), noiseless=True)
    np.testing.assert_allclose(mean2, mean, rtol=1e-9, atol=1e-10)


# ---------------------------------------------------------------------------------------------- prior programs
def _mean_fn(x, params):                          # tests/test_gp.py:25-26
    return params["a"] * x ** params["b"]


def _mean_fn_priors():                            # tests/test_gp.py:29-32, numpyro -> gpax_b200.priors
    from gpax_b200 import priors as numpyro
    a = numpyro.sample("a", numpyro.distributions.LogNormal(0, 1))
    b = numpyro.sample("b", numpyro.distributions.Normal(3, 1))
    return {"a": a, "b": b}


def _kernel_custom_prior():                       # tests/test_gp.py:35-38
    from gpax_b200 import priors as numpyro
    length = numpyro.sample("k_length", numpyro.distributions.Uniform(0, 1))
    scale = numpyro.sample("k_scale", numpyro.distributions.LogNormal(0, 1))
    return {"k_length": length, "k_scale": scale}


def _dummy_data():                                # tests/test_gp.py:15-22, seeded
    rng = np.random.default_rng(0)
    X = np.linspace(1, 2, 8) + 0.1 * rng.standard_normal(8)
    return X, 10 * X ** 2


def test_program_log_joint_gradient_through_the_gpu():
    """the log joint of gp.py:137-164 with kernel_prior and mean_fn_prior programs: GPU likelihood + alpha, host chain
    rule; checked against central differences of its own value and against the NumPy likelihood"""
    import gpax_b200
    from gpax_b200.inference import make_log_joint, ProgramLogJoint
    X, y = _dummy_data()
    with pytest.warns(UserWarning):
        m = gpax_b200.ExactGP(1, "Matern", mean_fn=_mean_fn, mean_fn_prior=_mean_fn_priors, kernel_prior=_kernel_custom_prior)
    m.X_train, m.y_train = m._set_data(X, y)
    lj = make_log_joint(m)
    assert isinstance(lj, ProgramLogJoint) and lj.dim == 5
    u = np.array([0.4, 0.3, -1.2, 2.2, 2.05])
    val, g = lj(u, True)
    h, fd = 1e-5, np.zeros(5)
    for k in range(5):
        e = np.zeros(5)
        e[k] = h
        fd[k] = (lj(u + e, True)[0] - lj(u - e, True)[0]) / (2 * h)
    np.testing.assert_allclose(g, fd, rtol=1e-5, atol=1e-6)
    th, mean, _ = lj._run(u)
    from gpax_b200 import priors as P
    want = mll_ref("Matern", X[:, None], y - mean, th, 1e-6)
    want += float(P.Uniform(0, 1).log_prob(th[0])) + float(P.LogNormal().log_prob(th[1])) + float(P.LogNormal().log_prob(th[2]))
    want += float(P.LogNormal().log_prob(np.exp(2.2))) + float(P.Normal(3, 1).log_prob(2.05))
    s = 1 / (1 + np.exp(-0.4))
    want += np.log(s * (1 - s)) + 0.3 - 1.2 + 2.2
    assert abs(val - want) < 1e-8 * max(1.0, abs(want))


@pytest.mark.parametrize("kernel", ["RBF", "Matern"])
def test_fit_with_custom_kernel_priors(kernel):
    """tests/test_gp.py:129-136"""
    import gpax_b200
    X, y = _dummy_data()
    with pytest.warns(UserWarning):
        m = gpax_b200.ExactGP(1, kernel, kernel_prior=_kernel_custom_prior)
    m.fit(0, X, y, num_warmup=50, num_samples=50, progress_bar=False, print_summary=False)
    s = m.get_samples()
    assert m.mcmc is not None and s["k_length"].shape == (50,) and ((s["k_length"] > 0) & (s["k_length"] < 1)).all()


def test_fit_predict_with_prob_mean_fn():
    """tests/test_gp.py:294-301, 317-330"""
    import gpax_b200
    X, y = _dummy_data()
    m = gpax_b200.ExactGP(1, "RBF", mean_fn=_mean_fn, mean_fn_prior=_mean_fn_priors)
    m.fit(0, X, y, num_warmup=100, num_samples=100, progress_bar=False, print_summary=False)
    s = m.get_samples()
    assert set(s) == {"k_length", "k_scale", "noise", "a", "b"} and s["a"].shape == (100,)
    y_pred, y_sampled = m.predict(1, X)
    assert y_pred.shape == X.shape and y_sampled.shape == (100, 1, X.shape[0])
    assert np.abs(y_pred - y).max() < 0.15 * np.abs(y).max()


def test_vigp_fit_with_prob_mean_fn():
    """tests/test_vigp.py: SVI with a probabilistic mean function (delta guide median carries the mean parameters)"""
    import gpax_b200
    X, y = _dummy_data()
    m = gpax_b200.viGP(1, "RBF", mean_fn=_mean_fn, mean_fn_prior=_mean_fn_priors)
    m.fit(0, X, y, num_steps=200, step_size=0.05, progress_bar=False, print_summary=False)
    assert {"a", "b"} <= set(m.kernel_params)
    mean, var = m.predict(0, X)
    assert mean.shape == X.shape and (var > 0).all() and m.svi.losses[-10:].mean() < m.svi.losses[:10].mean()


def test_sample_from_prior():
    """tests/test_gp.py:333-338 contract + the draws have the prior-predictive variance"""
    import gpax_b200
    from gpax_b200 import priors as P
    X, _ = _dummy_data()
    m = gpax_b200.ExactGP(1, "RBF")
    y = m.sample_from_prior(0, X, num_samples=8)
    assert y.shape == (8, X.shape[0]) and np.isfinite(y).all()
    np.testing.assert_array_equal(y, m.sample_from_prior(0, X, num_samples=8))       # same key, same draws
    # narrow priors: Var y_i = k_scale + noise + jitter, Cov(y_i, y_j) = k_scale exp(-0.5 (x_i - x_j)^2 / l^2)
    tight = lambda v: P.LogNormal(np.log(v), 1e-3)                                     # noqa: E731
    def kp():
        return {"k_length": P.sample("k_length", tight(0.5)), "k_scale": P.sample("k_scale", tight(2.0))}
    with pytest.warns(UserWarning):
        m2 = gpax_b200.ExactGP(1, "RBF", kernel_prior=kp, noise_prior_dist=tight(0.3), mean_fn=lambda x: 3.0 * x.squeeze())
    Xg = np.array([0.0, 0.25, 2.0])
    ys = m2.sample_from_prior(1, Xg, num_samples=6000)
    C = np.cov(ys.T)
    want = 2.0 * np.exp(-0.5 * (Xg[:, None] - Xg[None]) ** 2 / 0.25) + 0.3 * np.eye(3)
    assert np.abs(C - want).max() < 0.15 and np.abs(ys.mean(0) - 3.0 * Xg).max() < 0.08
