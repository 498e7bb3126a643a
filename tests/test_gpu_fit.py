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


@pytest.mark.parametrize("kind,d,N", [("RBF", 1, 50), ("Matern", 2, 300), ("Periodic", 1, 130), ("RBF", 3, 700)])
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
    mb, vb = m.predict_in_batches(0, Xt, 10, params)
    assert fn(m.ctx.h) - h0 >= 10
    np.testing.assert_allclose(mb, mean, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(vb, var, rtol=1e-11, atol=1e-13)
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
