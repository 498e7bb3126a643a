"""CPU: host-side logic of the shell (no GPU calls): parameter packing, chunking, kernel table, seeds."""
import numpy as np
import pytest

import gpax_b200
from gpax_b200.gp import _theta_rows
from gpax_b200.utils import get_keys, initialize_inducing_points, seed_from_key, split_in_batches


def test_theta_rows_single_and_batched():
    th = _theta_rows({"k_length": np.array([0.5]), "k_scale": 2.0, "noise": 0.1}, 3, False)
    assert th.shape == (1, 6)
    np.testing.assert_array_equal(th[0], [0.5, 0.5, 0.5, 2.0, 0.1, 1.0])
    th = _theta_rows({"k_length": np.array([0.5, 0.6]), "k_scale": np.array(2.0), "noise": np.array(0.1),
                      "period": None}, 2, False)
    np.testing.assert_array_equal(th[0], [0.5, 0.6, 2.0, 0.1, 1.0])
    S = 7
    rng = np.random.default_rng(0)
    samples = {"k_length": rng.uniform(1, 2, (S, 1)), "k_scale": rng.uniform(1, 2, S), "noise": rng.uniform(0, 1, S),
               "period": rng.uniform(1, 2, S)}
    th = _theta_rows(samples, 2, True)
    assert th.shape == (S, 5)
    np.testing.assert_array_equal(th[:, 0], samples["k_length"][:, 0])
    np.testing.assert_array_equal(th[:, 1], samples["k_length"][:, 0])
    np.testing.assert_array_equal(th[:, 4], samples["period"])
    with pytest.raises(ValueError):
        _theta_rows({"k_length": np.ones(3), "k_scale": 1.0, "noise": 0.1}, 2, False)


def test_split_in_batches_matches_reference_lengths(golden):
    A = np.arange(23.0)[:, None]
    for bs in (2, 3, 8, 23):
        parts = split_in_batches(A, bs)
        assert [len(p) for p in parts] == list(golden[f"split23_bs{bs}_lens"])
    # fewer rows than batch_size: the reference raises UnboundLocalError; here one short chunk
    assert [len(p) for p in split_in_batches(A, 100)] == [23]
    assert [p.shape for p in split_in_batches(np.zeros((3, 10)), 4, dim=1)] == [(3, 4), (3, 4), (3, 2)]
    with pytest.raises(NotImplementedError):
        split_in_batches(A, 2, dim=2)


def test_get_kernel_table():
    assert gpax_b200.get_kernel("RBF") is gpax_b200.RBFKernel
    assert gpax_b200.get_kernel("Matern") is gpax_b200.MaternKernel
    assert gpax_b200.get_kernel("Periodic") is gpax_b200.PeriodicKernel
    f = lambda X, Z, p, n=0, **kw: None   # noqa: E731
    assert gpax_b200.get_kernel(f) is f
    with pytest.raises(KeyError):
        gpax_b200.get_kernel("NoSuchKernel")


def test_constructor_surface():
    m = gpax_b200.ExactGP(2, "Matern")
    assert m.kernel_dim == 2 and m.kernel_name == "Matern" and m.X_train is None and m.mcmc is None
    X, y = m._set_data(np.arange(5.0), np.arange(5.0)[:, None])
    assert X.shape == (5, 1) and y.shape == (5,)
    v = gpax_b200.viGP(1, "RBF", guide="normal")
    assert v.guide_type == "normal" and v.svi is None
    s = gpax_b200.viSparseGP(1, "RBF")
    assert s.Xu is None
    with pytest.warns(FutureWarning):
        gpax_b200.ExactGP(1, "RBF", noise_prior=lambda: None)


def test_keys_and_seeds():
    k1, k2 = get_keys(3)
    assert k1.dtype == np.uint32 and k1.shape == (2,) and not np.array_equal(k1, k2)
    a = seed_from_key(k1).standard_normal(4)
    b = seed_from_key(k1).standard_normal(4)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, seed_from_key(k2).standard_normal(4))
    np.testing.assert_array_equal(seed_from_key(5).standard_normal(3), seed_from_key(np.array(5)).standard_normal(3))


def test_inducing_points():
    X = np.arange(400.0).reshape(200, 2)
    u = initialize_inducing_points(X, 0.1, "uniform")
    assert u.shape == (20, 2) and np.array_equal(u[0], X[0]) and np.array_equal(u[-1], X[-1])
    r = initialize_inducing_points(X, 0.25, "random", key=1)
    assert r.shape == (50, 2) and len({tuple(v) for v in r}) == 50
    with pytest.raises(ValueError):
        initialize_inducing_points(X, 1.5)
