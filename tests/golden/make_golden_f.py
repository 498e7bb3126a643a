"""
make_golden_f.py -- golden vectors for the SURVEY.md section 8f rows (model variants, acquisition functions), generated
by executing the REFERENCE'S OWN SOURCE (/root/reference/gpax, unmodified) over the same NumPy shim of jax.numpy as
make_golden.py, plus three functional stand-ins the 8f code paths need at run time:

  * numpyro.distributions.Normal(loc, scale).cdf / .log_prob        -> scipy.special.ndtr / the closed form
  * numpyro.distributions.MultivariateNormal(mean, cov).sample(key, (n,)) -> mean + eps @ chol(cov)^T with the eps
    array held in `INJECT["eps"]` (JAX's threefry stream is pinned separately, tests/test_prng.py)
  * jax.vmap with in_axes / dict arguments / keyword arguments      -> Python loops

What runs for real: gpax.acquisition.base_acq.ei / ucb / ue / poi / kg, gpax.models.hskgp.VarNoiseGP.get_mvn_posterior,
gpax.models.vgp.vExactGP.get_mvn_posterior, gpax.models.uigp.UIGP.get_mvn_posterior, gpax.kernels.NNGPKernel,
gpax.models.mngp.MeasuredNoiseGP's covariance k + diag(measured_noise) (mngp.py:92-97).

Run:  python tests/golden/make_golden_f.py     (needs /root/reference; writes tests/golden/reference_vectors_f.npz)
"""
import os
import sys
import types

import numpy as np
import scipy.linalg
import scipy.special

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg   # noqa: E402

JArr = mg.JArr
INJECT = {"eps": None}


def _tree_index(a, i, axis):
    if axis is None or a is None:
        return a
    if isinstance(axis, dict):
        return {k: _tree_index(a[k], i, axis[k]) for k in a}
    if isinstance(a, dict):
        return {k: _tree_index(v, i, axis) for k, v in a.items()}
    if isinstance(a, (tuple, list)):
        return type(a)(_tree_index(v, i, axis) for v in a)
    return np.asarray(a)[i].view(JArr) if np.ndim(a) > 0 else a


def _tree_len(a):
    if isinstance(a, dict):
        return _tree_len(next(iter(a.values())))
    if isinstance(a, (tuple, list)):
        return _tree_len(a[0])
    return len(a)


def vmap(f, in_axes=0, out_axes=0):
    def g(*args, **kwargs):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if isinstance(ax, dict):
                n = len(next(a[k] for k in a if ax[k] is not None))
                break
            if ax is not None and a is not None:
                n = _tree_len(a)
                break
        outs = []
        for i in range(n):
            ai = [_tree_index(a, i, ax) for a, ax in zip(args, axes)]
            ki = {k: _tree_index(v, i, 0) for k, v in kwargs.items()}
            outs.append(f(*ai, **ki))
        if isinstance(outs[0], tuple):
            return tuple(np.stack([np.asarray(o[j]) for o in outs]).view(JArr) for j in range(len(outs[0])))
        return np.stack([np.asarray(o) for o in outs]).view(JArr)
    return g


class Normal:
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = np.asarray(loc, dtype=np.float64), np.asarray(scale, dtype=np.float64)

    def cdf(self, v):
        return scipy.special.ndtr((np.asarray(v) - self.loc) / self.scale).view(JArr)

    def log_prob(self, v):
        z = (np.asarray(v) - self.loc) / self.scale
        return (-0.5 * z * z - np.log(self.scale) - 0.5 * np.log(2 * np.pi)).view(JArr)


class MultivariateNormal:
    def __init__(self, loc=None, covariance_matrix=None, **kw):
        self.loc, self.cov = np.asarray(loc), np.asarray(covariance_matrix)

    def sample(self, key, sample_shape=()):
        eps = INJECT["eps"]
        L = np.linalg.cholesky(self.cov)
        return (self.loc[None, :] + eps @ L.T).view(JArr)


def install():
    mg.install_shim()
    import jax
    jax.vmap = vmap
    sys.modules["jax"].vmap = vmap
    dist = sys.modules.get("numpyro.distributions")
    if dist is None:
        import numpyro.distributions as dist   # noqa: F401  (stub module created by the finder)
        dist = sys.modules["numpyro.distributions"]
    dist.Normal = Normal
    dist.MultivariateNormal = MultivariateNormal


def J(a):
    return np.asarray(a, dtype=np.float64).view(JArr)


def main():
    if not os.path.isdir(mg.REF):
        raise SystemExit("reference tree not present; golden vectors are generated in the build container only")
    install()
    sys.path.insert(0, mg.REF)
    import gpax
    from gpax.acquisition import base_acq
    import gpax.kernels.kernels as rk
    import gpax.kernels.mtkernels as rmt
    rk.vmap = vmap                      # kernels.py does `from jax import vmap`
    rmt.vmap = vmap
    base_acq.jax.vmap = vmap
    rng = np.random.default_rng(20260925)
    out = {}

    # ---- base acquisition functions on moments (base_acq.py:20-155)
    P = 37
    mean = rng.standard_normal(P) * 2.0
    var = np.exp(rng.normal(-1.0, 1.0, P))
    var[5] = 1e-12
    out["acq_mean"], out["acq_var"] = mean, var
    for mx in (False, True):
        for bf in (None, 0.3):
            tag = f"mx{int(mx)}_bf{'none' if bf is None else 'given'}"
            out[f"acq_ei_{tag}"] = np.asarray(base_acq.ei((J(mean), J(var)), bf, mx))
            out[f"acq_poi_{tag}"] = np.asarray(base_acq.poi((J(mean), J(var)), bf, 0.01, mx))
        out[f"acq_ucb_mx{int(mx)}"] = np.asarray(base_acq.ucb((J(mean), J(var)), 0.25, mx))
    out["acq_ue"] = np.asarray(base_acq.ue((J(mean), J(var))))

    # ---- knowledge gradient (base_acq.py:158-232): the reference's literal re-inversion per candidate and simulation
    N, Pk, nsim = 40, 12, 5
    Xtr = rng.uniform(0, 1, (N, 2))
    ytr = np.sin(5 * Xtr[:, 0]) + Xtr[:, 1] ** 2 + 0.05 * rng.standard_normal(N)
    Xc = rng.uniform(0, 1, (Pk, 2))
    params = {"k_length": J([0.4, 0.5]), "k_scale": np.float64(1.3), "noise": np.float64(0.05)}
    eps = rng.standard_normal((nsim, Pk))
    out["kg_Xtr"], out["kg_ytr"], out["kg_Xc"], out["kg_eps"] = Xtr, ytr, Xc, eps
    for kname in ("RBF", "Matern"):
        for mx in (True, False):
            for nl in (True, False):
                m = gpax.ExactGP(2, kernel=kname)
                m.X_train, m.y_train = J(Xtr), J(ytr)
                INJECT["eps"] = eps
                v = base_acq.kg(m, J(Xc), params, rng_key=np.zeros(2, np.uint32), n=nsim, maximize=mx, noiseless=nl)
                out[f"kg_{kname}_mx{int(mx)}_nl{int(nl)}"] = np.asarray(v)

    # ---- VarNoiseGP.get_mvn_posterior (hskgp.py:165-206)
    N, Pn = 60, 15
    Xtr = rng.uniform(0, 1, (N, 1))
    ytr = np.sin(6 * Xtr[:, 0]) + 0.1 * rng.standard_normal(N)
    Xte = np.linspace(0, 1, Pn)[:, None]
    log_var = rng.normal(-3.0, 0.4, N)
    params = {"k_length": J([0.3]), "k_scale": np.float64(1.1), "k_noise_length": J([0.5]),
              "k_noise_scale": np.float64(0.7), "log_var": J(log_var), "noise": np.float64(0.0)}
    m = gpax.VarNoiseGP(1, kernel="RBF", noise_kernel="Matern")
    m.X_train, m.y_train = J(Xtr), J(ytr)
    mean, cov = m.get_mvn_posterior(J(Xte), params)
    out["hsk_Xtr"], out["hsk_ytr"], out["hsk_Xte"], out["hsk_log_var"] = Xtr, ytr, Xte, log_var
    out["hsk_mean"], out["hsk_cov"] = np.asarray(mean), np.asarray(cov)

    # ---- vExactGP.get_mvn_posterior (vgp.py:125-172): outer task batch
    B, N, Pv = 3, 50, 11
    Xtr = rng.uniform(0, 1, (B, N, 2))
    ytr = np.stack([np.sin((3 + b) * Xtr[b, :, 0]) * np.cos(2 * Xtr[b, :, 1]) + 0.05 * rng.standard_normal(N) for b in range(B)])
    Xte = rng.uniform(0, 1, (B, Pv, 2))
    params = {"k_length": J(rng.uniform(0.3, 0.6, (B, 2))), "k_scale": J(rng.uniform(0.8, 1.5, B)),
              "noise": J(rng.uniform(0.02, 0.1, B))}
    m = gpax.vExactGP(2, kernel="Matern")
    m.X_train, m.y_train = J(Xtr), J(ytr)
    mean, cov = m.get_mvn_posterior(J(Xte), params, noiseless=False)
    out["vgp_Xtr"], out["vgp_ytr"], out["vgp_Xte"] = Xtr, ytr, Xte
    for k in ("k_length", "k_scale", "noise"):
        out["vgp_" + k] = np.asarray(params[k])
    out["vgp_mean"], out["vgp_cov"] = np.asarray(mean), np.asarray(cov)

    # ---- UIGP.get_mvn_posterior (uigp.py:131-150): the training inputs are a per-draw parameter X_prime
    N, Pu = 45, 9
    Xtr = rng.uniform(0, 1, (N, 1))
    ytr = np.cos(5 * Xtr[:, 0]) + 0.05 * rng.standard_normal(N)
    Xprime = Xtr + 0.03 * rng.standard_normal((N, 1))
    Xte = np.linspace(0, 1, Pu)[:, None]
    params = {"k_length": J([0.35]), "k_scale": np.float64(0.9), "noise": np.float64(0.04), "X_prime": J(Xprime)}
    m = gpax.UIGP(1, kernel="RBF")
    m.X_train, m.y_train = J(Xtr), J(ytr)
    mean, cov = m.get_mvn_posterior(J(Xte), params, noiseless=True)
    out["uigp_Xtr"], out["uigp_ytr"], out["uigp_Xprime"], out["uigp_Xte"] = Xtr, ytr, Xprime, Xte
    out["uigp_mean"], out["uigp_cov"] = np.asarray(mean), np.asarray(cov)

    # ---- NNGP kernel (kernels.py:120-224)
    X = rng.standard_normal((9, 3))
    Z = rng.standard_normal((6, 3))
    out["nngp_X"], out["nngp_Z"] = X, Z
    for act in ("erf", "relu"):
        for depth in (1, 3):
            kfn = gpax.kernels.NNGPKernel(activation=act, depth=depth)
            prm = {"var_b": np.float64(0.3), "var_w": np.float64(1.7)}
            out[f"nngp_{act}_d{depth}_XZ"] = np.asarray(kfn(J(X), J(Z), prm, 0.05))
            out[f"nngp_{act}_d{depth}_XX"] = np.asarray(kfn(J(X), J(X), prm, 0.05))

    # ---- multi-task kernels (mtkernels.py:61-232)
    Tn = 3
    W = rng.standard_normal((Tn, 2))
    v = np.exp(rng.normal(-1, 0.3, Tn))
    Xm = np.column_stack([rng.uniform(0, 1, (14, 2)), rng.integers(0, Tn, 14)])
    Zm = np.column_stack([rng.uniform(0, 1, (9, 2)), rng.integers(0, Tn, 9)])
    prm = {"k_length": J([0.4, 0.6]), "k_scale": np.float64(1.2), "W": J(W), "v": J(v)}
    noise_t = J([0.01, 0.02, 0.03])
    kmt = gpax.kernels.MultitaskKernel("Matern")
    out["mt_X"], out["mt_Z"], out["mt_W"], out["mt_v"] = Xm, Zm, W, v
    out["mt_XZ"] = np.asarray(kmt(J(Xm), J(Zm), prm, noise_t))
    out["mt_XX"] = np.asarray(kmt(J(Xm), J(Xm), prm, noise_t))
    # (a scalar noise becomes a length-1 array indexed by the task ids, mtkernels.py:113-115: that relies on JAX clamping
    #  out-of-range gathers, which NumPy does not do -- not generated here)
    kmv = gpax.kernels.MultivariateKernel("RBF", Tn)
    out["mv_XZ"] = np.asarray(kmv(J(Xm[:, :2]), J(Zm[:, :2]), prm, noise_t))
    out["mv_XX"] = np.asarray(kmv(J(Xm[:, :2]), J(Xm[:, :2]), prm, noise_t))
    klcm = gpax.kernels.LCMKernel("RBF", shared_input_space=False)
    prm2 = {"k_length": J(rng.uniform(0.3, 0.7, (2, 2))), "k_scale": J([1.1, 0.7]), "W": J(rng.standard_normal((2, Tn, 2))),
            "v": J(np.exp(rng.normal(-1, 0.3, (2, Tn))))}
    for k_ in ("k_length", "k_scale", "W", "v"):
        out["lcm_" + k_] = np.asarray(prm2[k_])
    out["lcm_XX"] = np.asarray(klcm(J(Xm), J(Xm), prm2, noise_t))

    # ---- MeasuredNoiseGP: the model covariance k + diag(measured_noise) (mngp.py:92-97) and the log density of y under it
    N = 30
    Xtr = rng.uniform(0, 1, (N, 1))
    ytr = np.sin(4 * Xtr[:, 0]) + 0.1 * rng.standard_normal(N)
    mnoise = np.exp(rng.normal(-3.5, 0.5, N))
    prm = {"k_length": J([0.4]), "k_scale": np.float64(1.2)}
    k = np.asarray(gpax.kernels.RBFKernel(J(Xtr), J(Xtr), prm, 0, jitter=1e-6)) + np.diag(mnoise)
    Lk = scipy.linalg.cholesky(k, lower=True)
    a = scipy.linalg.solve_triangular(Lk, ytr, lower=True)
    out["mn_Xtr"], out["mn_ytr"], out["mn_noise"], out["mn_cov"] = Xtr, ytr, mnoise, k
    out["mn_logp"] = np.array(-0.5 * a @ a - np.log(np.diag(Lk)).sum() - 0.5 * N * np.log(2 * np.pi))

    np.savez_compressed(os.path.join(HERE, "reference_vectors_f.npz"), **out)
    print("wrote reference_vectors_f.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
