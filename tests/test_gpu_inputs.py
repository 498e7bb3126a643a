"""GPU: input handling of the shell (jnp or NumPy, 1-D or (n,1), float32 default in the reference)."""
import numpy as np
import pytest

import oracle
from conftest import assert_close

pytestmark = pytest.mark.gpu


def test_float32_noncontiguous_and_1d_inputs():
    import gpax_b200
    rng = np.random.default_rng(0)
    Xbig = rng.uniform(0, 1, (200, 4))
    X = Xbig[::2, 1:3]                       # non-contiguous view, d = 2
    y = np.sin(4 * X[:, 0]) + X[:, 1]
    Xt = rng.uniform(0, 1, (30, 2))
    params = {"k_length": np.array([0.5, 0.6], dtype=np.float32), "k_scale": np.float32(1.0), "noise": np.float32(0.1)}
    m = gpax_b200.ExactGP(2, "Matern")
    m.X_train, m.y_train = X.astype(np.float32), y.astype(np.float32)[:, None]      # (n,1) targets, float32
    mean, cov = m.get_mvn_posterior(Xt.astype(np.float32), params)
    assert mean.dtype == np.float32 and cov.dtype == np.float32
    p64 = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    rm, rc = oracle.exact_posterior(X.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64),
                                    Xt.astype(np.float32).astype(np.float64), p64, "Matern")
    np.testing.assert_allclose(mean, rm, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cov, rc, rtol=1e-4, atol=1e-5)
    # python lists and 1-D inputs
    m1 = gpax_b200.viGP(1, "RBF")
    m1.X_train, m1.y_train = list(np.linspace(0, 1, 20)), list(np.linspace(0, 1, 20) ** 2)
    mu, var = m1.predict(0, [0.1, 0.5, 0.9], {"k_length": 0.3, "k_scale": 1.0, "noise": 0.01})
    assert mu.shape == (3,) and var.shape == (3,)
    rm, rv = oracle.vi_predict(np.linspace(0, 1, 20), np.linspace(0, 1, 20) ** 2, np.array([0.1, 0.5, 0.9]),
                               {"k_length": np.array([0.3]), "k_scale": 1.0, "noise": 0.01}, "RBF")
    assert_close(mu, rm, 1e-9)
    assert_close(var, rv, 1e-7)


def test_bad_arguments_raise():
    import gpax_b200
    from gpax_b200 import B200GPError
    ctx = gpax_b200.default_context()
    with pytest.raises(ValueError):
        gpax_b200.RBFKernel(np.zeros((3, 2)), np.zeros((3, 3)), {"k_length": 1.0, "k_scale": 1.0})
    with pytest.raises(B200GPError):
        ctx.set_option("no_such_option", 1)
    with pytest.raises(B200GPError):
        ctx.gram("RBF", np.zeros((3, 70)), np.zeros((3, 70)), np.ones(70), 1.0)      # d > 64 unsupported
    m = gpax_b200.ExactGP(2, "RBF")
    m.X_train, m.y_train = np.zeros((5, 2)), np.zeros(5)
    with pytest.raises(ValueError):
        m.get_mvn_posterior(np.zeros((3, 2)), {"k_length": np.ones(3), "k_scale": 1.0, "noise": 0.1})


def test_fp32_io_matches_fp32_rounded_oracle():
    """B2GP_FLAG_F32 (the reference's default precision, gpax/utils/utils.py:19-21): float32 arrays in and out of the C-ABI,
    fp64 in between -- the result is the fp64 posterior of the float32-rounded inputs, rounded once to float32."""
    import gpax_b200
    import oracle
    rng = np.random.default_rng(12)
    N, P, d = 900, 70, 2
    X = rng.uniform(0, 1, (N, d)).astype(np.float32)
    y = (np.sin(5 * X[:, 0]) + X[:, 1]).astype(np.float32)
    Xn = rng.uniform(0, 1, (P, d)).astype(np.float32)
    params = {"k_length": np.array([0.3, 0.4]), "k_scale": 1.2, "noise": 0.05}
    for kname in ("RBF", "Matern"):
        rmean, rcov = oracle.exact_posterior_chol(X.astype(np.float64), y.astype(np.float64), Xn.astype(np.float64), params, kname)
        m = gpax_b200.ExactGP(d, kname)
        m.X_train, m.y_train = X, y
        mean, cov = m.get_mvn_posterior(Xn, params)
        assert mean.dtype == np.float32 and cov.dtype == np.float32
        np.testing.assert_array_equal(mean, rmean.astype(np.float32))
        np.testing.assert_allclose(cov, rcov.astype(np.float32), rtol=0, atol=6e-8 * np.abs(rcov).max())   # one float32 rounding of entries that cancel
        K = gpax_b200.get_kernel(kname)(X, X, params, 0.05)
        assert K.dtype == np.float32
        ref = oracle.get_kernel(kname)(X.astype(np.float64), X.astype(np.float64), params, 0.05).astype(np.float32)
        np.testing.assert_array_equal(K, ref)
    v = gpax_b200.viGP(d, "RBF")
    v.X_train, v.y_train = X, y
    vm, vv = v.predict(None, Xn, samples=params)
    assert vm.dtype == np.float32 and vv.dtype == np.float32
    ctx = m.ctx
    out = ctx.posterior("RBF", X, y, Xn, np.array([[0.3, 0.4, 1.2, 0.05, 1.0]]), want=("mean", "var"), f32=True)
    np.testing.assert_array_equal(out["mean"][0], vm)
