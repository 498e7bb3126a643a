"""CPU: the restatements of the section-8f rows (oracle/acq_oracle.py, oracle/variants_oracle.py) against the golden
vectors produced by the reference's own source (tests/golden/make_golden_f.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT, assert_close
from oracle import acq_oracle as ao
from oracle import variants_oracle as vo


@pytest.fixture(scope="module")
def gf():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_f.npz"))


def test_base_acquisition_functions(gf):
    mean, var = gf["acq_mean"], gf["acq_var"]
    for mx in (False, True):
        for bf, tag in ((None, "none"), (0.3, "given")):
            np.testing.assert_allclose(ao.ei(mean, var, bf, mx), gf[f"acq_ei_mx{int(mx)}_bf{tag}"], rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(ao.poi(mean, var, bf, 0.01, mx), gf[f"acq_poi_mx{int(mx)}_bf{tag}"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(ao.ucb(mean, var, 0.25, mx), gf[f"acq_ucb_mx{int(mx)}"], rtol=1e-14)
    np.testing.assert_allclose(ao.ue(mean, var), gf["acq_ue"], rtol=1e-15)


@pytest.mark.parametrize("kname", ["RBF", "Matern"])
def test_knowledge_gradient_literal(gf, kname):
    params = {"k_length": np.array([0.4, 0.5]), "k_scale": 1.3, "noise": 0.05}
    for mx in (True, False):
        for nl in (True, False):
            v = ao.kg(gf["kg_Xtr"], gf["kg_ytr"], gf["kg_Xc"], params, kname, gf["kg_eps"], mx, nl)
            assert_close(v, gf[f"kg_{kname}_mx{int(mx)}_nl{int(nl)}"], 1e-8, f"kg {kname} maximize={mx} noiseless={nl}")


def test_var_noise_gp(gf):
    params = {"k_length": np.array([0.3]), "k_scale": 1.1, "k_noise_length": np.array([0.5]), "k_noise_scale": 0.7,
              "log_var": gf["hsk_log_var"], "noise": 0.0}
    mean, cov = vo.var_noise_posterior(gf["hsk_Xtr"], gf["hsk_ytr"], gf["hsk_Xte"], params, "RBF", "Matern")
    assert_close(mean, gf["hsk_mean"], 1e-7)      # k_XX carries jitter only (cond ~ 1e7): the LU inverse itself is the limit
    assert_close(cov, gf["hsk_cov"], 1e-7)


def test_task_batch(gf):
    params = {k: gf["vgp_" + k] for k in ("k_length", "k_scale", "noise")}
    mean, cov = vo.task_batch_posterior(gf["vgp_Xtr"], gf["vgp_ytr"], gf["vgp_Xte"], params, "Matern")
    assert_close(mean, gf["vgp_mean"], 1e-10)
    assert_close(cov, gf["vgp_cov"], 1e-10)


def test_uigp(gf):
    params = {"k_length": np.array([0.35]), "k_scale": 0.9, "noise": 0.04, "X_prime": gf["uigp_Xprime"]}
    mean, cov = vo.uigp_posterior(gf["uigp_ytr"], gf["uigp_Xte"], params, "RBF", noiseless=True)
    assert_close(mean, gf["uigp_mean"], 1e-10)
    assert_close(cov, gf["uigp_cov"], 1e-10)


def test_nngp_kernel(gf):
    prm = {"var_b": 0.3, "var_w": 1.7}
    for act in ("erf", "relu"):
        for depth in (1, 3):
            np.testing.assert_allclose(vo.nngp_kernel(gf["nngp_X"], gf["nngp_Z"], prm, 0.05, activation=act, depth=depth),
                                       gf[f"nngp_{act}_d{depth}_XZ"], rtol=1e-12)
            np.testing.assert_allclose(vo.nngp_kernel(gf["nngp_X"], gf["nngp_X"], prm, 0.05, activation=act, depth=depth),
                                       gf[f"nngp_{act}_d{depth}_XX"], rtol=1e-12)


def test_measured_noise_logp(gf):
    prm = {"k_length": np.array([0.4]), "k_scale": 1.2}
    v = vo.measured_noise_logp(gf["mn_Xtr"], gf["mn_ytr"], prm, gf["mn_noise"])
    np.testing.assert_allclose(v, gf["mn_logp"], rtol=1e-12)


def test_multitask_kernels(gf):
    """oracle/variants_oracle.multitask_kernel / multivariate_kernel / lcm_kernel vs the reference's mtkernels.py"""
    prm = {"k_length": np.array([0.4, 0.6]), "k_scale": 1.2, "W": gf["mt_W"], "v": gf["mt_v"]}
    nt = np.array([0.01, 0.02, 0.03])
    np.testing.assert_allclose(vo.multitask_kernel(gf["mt_X"], gf["mt_Z"], prm, nt, "Matern"), gf["mt_XZ"], rtol=1e-13)
    np.testing.assert_allclose(vo.multitask_kernel(gf["mt_X"], gf["mt_X"], prm, nt, "Matern"), gf["mt_XX"], rtol=1e-13)
    np.testing.assert_allclose(vo.multivariate_kernel(gf["mt_X"][:, :2], gf["mt_Z"][:, :2], prm, nt, "RBF", 3), gf["mv_XZ"], rtol=1e-13)
    np.testing.assert_allclose(vo.multivariate_kernel(gf["mt_X"][:, :2], gf["mt_X"][:, :2], prm, nt, "RBF", 3), gf["mv_XX"], rtol=1e-13)
    prm2 = {k: gf["lcm_" + k] for k in ("k_length", "k_scale", "W", "v")}
    np.testing.assert_allclose(vo.lcm_kernel(gf["mt_X"], gf["mt_X"], prm2, nt, "RBF"), gf["lcm_XX"], rtol=1e-13)
