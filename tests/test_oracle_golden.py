"""CPU: the oracle (oracle/gp_oracle.py) against the golden vectors produced by running the
reference's own source (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import oracle
from conftest import assert_close

KMAP = {"RBF": oracle.rbf_kernel, "Matern": oracle.matern_kernel, "Periodic": oracle.periodic_kernel}


def test_gram_cases(golden):
    cases = [str(c) for c in golden["gram_cases"]]
    assert len(cases) == 27
    for tag in cases:
        kname = tag.split("_")[1]
        X, Z, ell = golden[tag + "_X"], golden[tag + "_Z"], golden[tag + "_ell"]
        params = {"k_length": ell, "k_scale": 1.3, "period": 0.9}
        K = KMAP[kname](X, Z, params, 0.05, jitter=1e-6)
        np.testing.assert_allclose(K, golden[tag + "_K"], rtol=1e-14, atol=1e-15, err_msg=tag)


def test_gram_same_array(golden):
    X = golden["gramself_X"]
    params = {"k_length": np.array([0.4, 0.6]), "k_scale": 2.0, "period": 1.0}
    for kname, fn in KMAP.items():
        K = fn(X, X, params, 0.1, jitter=1e-6)
        np.testing.assert_allclose(K, golden[f"gramself_{kname}_K"], rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("kname", ["RBF", "Matern", "Periodic"])
def test_exact8(golden, kname):
    Xtr, ytr, Xte = golden["exact8_Xtr"], golden["exact8_ytr"], golden["exact8_Xte"]
    params = {"k_length": np.array([1.0]), "k_scale": 1.0, "noise": 0.1, "period": 1.0}
    for nl in (0, 1):
        mean, cov = oracle.exact_posterior(Xtr, ytr, Xte, params, kname, noiseless=bool(nl))
        # N=8 with unit lengthscale on [1,2]: cond(K) ~ 1e2..1e6 depending on the kernel; LAPACK is
        # deterministic here so the restatement reproduces the shimmed reference to rounding.
        assert_close(mean, golden[f"exact8_{kname}_nl{nl}_mean"], 1e-9, "mean")
        assert_close(cov, golden[f"exact8_{kname}_nl{nl}_cov"], 1e-9, "cov")
    mean, cov = oracle.exact_posterior(Xtr, ytr, Xte, params, kname, jitter=1e-5)
    assert_close(mean, golden[f"exact8_{kname}_jit1e-5_mean"], 1e-9)
    assert_close(cov, golden[f"exact8_{kname}_jit1e-5_cov"], 1e-9)


@pytest.mark.parametrize("kname,N", [("RBF", 300), ("Matern", 384), ("Periodic", 200)])
def test_exact_medium_and_vi(golden, kname, N):
    tag = f"exact_{kname}_N{N}"
    Xtr, ytr, Xte, ell = (golden[tag + s] for s in ("_Xtr", "_ytr", "_Xte", "_ell"))
    params = {"k_length": ell, "k_scale": 1.2, "noise": 0.1, "period": 0.8}
    mean, cov = oracle.exact_posterior(Xtr, ytr, Xte, params, kname)
    assert_close(mean, golden[tag + "_mean"], 1e-10)
    assert_close(cov, golden[tag + "_cov"], 1e-10)
    vm, vv = oracle.vi_predict(Xtr, ytr, Xte, params, kname, noiseless=True)
    assert_close(vm, golden[tag + "_vimean"], 1e-10)
    assert_close(vv, golden[tag + "_vivar"], 1e-10)
    # the Cholesky formulation of the same posterior agrees with the LU-inverse formulation
    mc, cc = oracle.exact_posterior_chol(Xtr, ytr, Xte, params, kname)
    assert_close(mc, golden[tag + "_mean"], 1e-9)
    assert_close(cc, golden[tag + "_cov"], 1e-9)
    mc, vc = oracle.exact_posterior_chol(Xtr, ytr, Xte, params, kname, noiseless=True, diag_only=True)
    assert_close(vc, golden[tag + "_vivar"], 1e-9)


def test_mean_fn(golden):
    Xtr, ytr, Xte = golden["meanfn_Xtr"], golden["meanfn_ytr"], golden["meanfn_Xte"]
    params = {"k_length": np.array([0.5]), "k_scale": 1.0, "noise": 0.05, "a": 9.0, "b": 0.5}
    mfn = lambda x, p: p["a"] * x[:, 0] ** 2 + p["b"]   # noqa: E731
    mean, cov = oracle.exact_posterior(Xtr, ytr, Xte, params, "RBF", mean_fn=mfn, mean_fn_takes_params=True)
    assert_close(mean, golden["meanfn_mean"], 1e-10)
    assert_close(cov, golden["meanfn_cov"], 1e-10)


@pytest.mark.parametrize("tag,kname", [("sparse50", "RBF"), ("sparse400", "Matern")])
def test_sparse(golden, tag, kname):
    Xtr, ytr, Xu, Xte = (golden[tag + s] for s in ("_Xtr", "_ytr", "_Xu", "_Xte"))
    d = Xtr.shape[1]
    params = {"k_length": np.full(d, 0.4), "k_scale": 1.0, "noise": 0.1}
    for nl in (0, 1):
        mean, cov = oracle.sparse_posterior(Xtr, ytr, Xu, Xte, params, kname, noiseless=bool(nl), jitter=1e-5)
        assert_close(mean, golden[f"{tag}_nl{nl}_mean"], 1e-9)
        assert_close(cov, golden[f"{tag}_nl{nl}_cov"], 1e-9)


def test_split_in_batches(golden):
    A = np.arange(23.0)[:, None]
    for bs in (2, 3, 8, 23):
        parts = oracle.split_in_batches(A, bs)
        assert [len(p) for p in parts] == list(golden[f"split23_bs{bs}_lens"])
        assert np.array_equal(np.concatenate(parts), A)


def test_closed_forms():
    """Analytic known answers: N=1 posterior, and k(x,x)."""
    params = {"k_length": np.array([0.7]), "k_scale": 1.5, "noise": 0.2}
    X = np.array([[0.3]])
    y = np.array([2.0])
    Xs = np.array([[0.3], [0.9]])
    mean, cov = oracle.exact_posterior(X, y, Xs, params, "RBF")
    kxx = 1.5 + 0.2 + 1e-6
    ks = 1.5 * np.exp(-0.5 * ((Xs[:, 0] - 0.3) / 0.7) ** 2)
    np.testing.assert_allclose(mean, ks * 2.0 / kxx, rtol=1e-13)
    np.testing.assert_allclose(np.diag(cov), 1.5 + 0.2 + 1e-6 - ks ** 2 / kxx, rtol=1e-12)
