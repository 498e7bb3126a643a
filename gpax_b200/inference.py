"""
inference.py -- hyper-parameter inference behind `.fit()`.

The reference hands `ExactGP.model` (gpax/models/gp.py:137-164) to NumPyro: NUTS/MCMC for ExactGP.fit
(gp.py:207-218) and SVI with Adam(b1=0.5) and an AutoDelta / AutoNormal guide for viGP.fit (vigp.py:108-120).
Here the model's log joint is evaluated directly: the likelihood term and its gradient come from the GPU
(b2gp_mll: Cholesky + solves + fused gradient reduction), the priors are the LogNormal(0,1) defaults of
gp.py:222-247 (or gpax_b200.priors objects), and the samplers / optimiser are small host-side NumPy loops.
Custom `kernel_prior` / `noise_prior` / `mean_fn_prior` *programs* are run on the host through gpax_b200.priors'
`sample` / `plate` primitives (ProgramLogJoint); a program written against NumPyro itself cannot be interpreted here.
"""
import math

import numpy as np

from . import priors as P
from .utils import seed_from_key


class LogJoint:
    """log p(y, theta) over the unconstrained vector u, with gradient; theta = (k_length[d], k_scale, noise[, period]).
    Default priors and `*_prior_dist` objects only; `make_log_joint` picks ProgramLogJoint when the model carries prior
    programs (kernel_prior / noise_prior / mean_fn_prior)."""

    def __init__(self, model, jitter=1e-6):
        if model._fused is None:
            raise NotImplementedError("fit() needs kernel 'RBF', 'Matern' or 'Periodic'")
        if model.kernel_prior is not None or model.noise_prior is not None or \
                (model.mean_fn is not None and model.mean_fn_prior is not None):
            raise NotImplementedError("prior programs go through ProgramLogJoint (inference.make_log_joint)")
        self.m, self.jitter = model, float(jitter)
        X, y = model._train_arrays()
        self.X, self.d = X, X.shape[1]
        self.y = y if model.mean_fn is None else y - np.asarray(model.mean_fn(X), dtype=np.float64).squeeze()
        self.kind = model._fused
        d = self.d
        lp = model.lengthscale_prior_dist or P.LogNormal(0.0, 1.0)       # gp.py:235-239
        npd = model.noise_prior_dist or P.LogNormal(0.0, 1.0)            # gp.py:222-227
        self.names = ["k_length"] * d + ["k_scale", "noise"]
        self.priors = [lp] * d + [P.LogNormal(0.0, 1.0), npd]            # gp.py:240
        self.idx = list(range(d + 2))                                     # position in the (d+3) theta vector
        if self.kind == "Periodic":                                       # gp.py:241-244
            self.names.append("period")
            self.priors.append(P.LogNormal(0.0, 1.0))
            self.idx.append(d + 2)
        for pr in self.priors:
            if not isinstance(pr, P.Prior):
                raise TypeError("priors must be gpax_b200.priors objects (numpyro distributions cannot be used here)")
        self.dim = len(self.priors)
        self.n_evals = 0

    def theta_of(self, u):
        th = np.ones(self.d + 3)
        for k, (pr, i) in enumerate(zip(self.priors, self.idx)):
            th[i] = pr.transform(u[k])
        return th

    def init_u(self):
        """init_to_median: the reference's NUTS uses init_to_median(num_samples=10) (gp.py:208); exact medians here"""
        return np.array([float(pr.inverse(pr.median())) for pr in self.priors])

    def __call__(self, u, jacobian):
        """log p(y | theta(u)) + sum log p(theta_k) [+ log |dtheta/du| if jacobian]; returns (value, grad_u)"""
        th = self.theta_of(u)
        val, g, info = self._lik(th)
        self.n_evals += 1
        if info != 0 or not np.isfinite(val):
            return -np.inf, np.zeros(self.dim)
        grad = np.zeros(self.dim)
        for k, (pr, i) in enumerate(zip(self.priors, self.idx)):
            t = th[i]
            dt = float(pr.dtheta_du(u[k]))
            val += float(pr.log_prob(t))
            grad[k] = g[i] / t * dt + float(pr.dlog_prob(t)) * dt       # g is d/dlog(theta)
            if jacobian:
                val += float(pr.log_abs_jac(u[k]))
                grad[k] += float(pr.dlog_abs_jac(u[k]))
        return val, grad

    def _lik(self, th):
        """log p(y | theta) and its gradient w.r.t. log(theta): the exact marginal likelihood (b2gp_mll)"""
        val, g, _, info = self.m.ctx.mll(self.kind, self.X, self.y, th, self.jitter, want_grad=True)
        return val, g, info

    def to_dict(self, U):
        """rows of unconstrained vectors -> dict of constrained samples with the reference's site names / shapes"""
        U = np.atleast_2d(U)
        th = np.stack([self.theta_of(u) for u in U])
        out = {"k_length": th[:, :self.d], "k_scale": th[:, self.d], "noise": th[:, self.d + 1]}
        if self.kind == "Periodic":
            out["period"] = th[:, self.d + 2]
        return out


def gp_model_program(m, d, kind):
    """the host side of ExactGP.model (gp.py:137-154): every statement except the likelihood.  Runs under
    priors.run_program; returns (kernel-parameter dict, noise, mean-function parameter dict or None)."""
    if m.kernel_prior is not None:
        kp = m.kernel_prior()
    else:                                                              # gp.py:229-247
        lp = m.lengthscale_prior_dist or P.LogNormal(0.0, 1.0)
        with P.plate("ard", d):
            length = P.sample("k_length", lp)
        kp = {"k_length": length, "k_scale": P.sample("k_scale", P.LogNormal(0.0, 1.0))}
        if kind == "Periodic":
            kp["period"] = P.sample("period", P.LogNormal(0.0, 1.0))
    if m.noise_prior is not None:
        noise = m.noise_prior()
    else:                                                              # gp.py:222-227
        noise = P.sample("noise", m.noise_prior_dist or P.LogNormal(0.0, 1.0))
    mp = m.mean_fn_prior() if (m.mean_fn is not None and m.mean_fn_prior is not None) else None
    return kp, noise, mp


def prior_draws(model, rng, num_samples, d):
    """`num_samples` runs of the model's prior statements with every site drawn from its prior (what NumPyro's Predictive
    does for gp.py:401-408): list of (kernel params, noise, mean params)"""
    kind = model.kernel_name if isinstance(model.kernel_name, str) else None
    return [P.run_program(lambda: gp_model_program(model, d, kind), rng=rng)[0] for _ in range(int(num_samples))]


class ProgramLogJoint:
    """The log joint of ExactGP.model (gp.py:137-164) when some of its priors are *programs*: `kernel_prior()` returning
    the kernel-parameter dict (gp.py:141-142), the deprecated `noise_prior()` (gp.py:146-147), `mean_fn_prior()` feeding a
    parametric mean function (gp.py:151-154).  The programs are written against gpax_b200.priors' `sample` / `plate`
    (the reference's run under NumPyro).  The model program is re-run on the host for every evaluation:

      u  --transform per site-->  site values  --programs-->  theta[d+3], mean vector m[N]
      value = log N(y - m; 0, K_theta) [GPU: b2gp_mll]  +  sum_sites log p(site)  (+ log |d site / du|)

    Gradient.  The GPU returns d value / d log(theta) and alpha = K^-1 (y - m) = d value / d m.  The programs are plain
    Python (no tracer to differentiate them), so the Jacobians d theta / du and d m / du are taken by central differences
    of the host program -- 2 dim runs of a function of a handful of scalars and one [N]-vector, exact to ~1e-10 relative;
    the N^3 part is never differenced.  Site log-densities are differentiated analytically; when a program makes one
    site's distribution depend on another site's value (a hierarchical prior), the prior term's gradient is differenced
    too."""

    FD_STEP = 1e-5

    def __init__(self, model, jitter=1e-6, lik=None):
        if model._fused is None:
            raise NotImplementedError("fit() needs kernel 'RBF', 'Matern' or 'Periodic'")
        self.m, self.jitter = model, float(jitter)
        X, y = model._train_arrays()
        self.X, self.d, self.y0 = X, X.shape[1], y
        self.kind = model._fused
        self.has_mean_params = model.mean_fn is not None and model.mean_fn_prior is not None
        self.fixed_mean = None
        if model.mean_fn is not None and not self.has_mean_params:
            self.fixed_mean = np.asarray(model.mean_fn(X), dtype=np.float64).squeeze()
        self._lik_fn = lik
        _, sites, _ = P.run_program(self._model_program)
        self.sites = list(sites.values())
        self.dim = sum(s.size for s in self.sites)
        self.n_evals = 0
        # does any site's distribution depend on the values of the others?  (two runs at different values)
        vals = {s.name: np.asarray(s.prior.transform(np.full(s.shape, 0.37))) for s in self.sites}
        _, sites2, _ = P.run_program(self._model_program, vals)
        self.hierarchical = any(vars(sites[k].prior) != vars(sites2[k].prior) for k in sites)

    def _model_program(self):
        return gp_model_program(self.m, self.d, self.kind)

    def _run(self, u):
        """u -> (theta[d+3], mean vector or None, sites with values)"""
        vals, o = {}, 0
        for s in self.sites:
            vals[s.name] = np.asarray(s.prior.transform(u[o:o + s.size])).reshape(s.shape)
            o += s.size
        (kp, noise, mp), sites, _ = P.run_program(self._model_program, vals)
        th = np.ones(self.d + 3)
        th[:self.d] = np.broadcast_to(np.asarray(kp["k_length"], dtype=np.float64).reshape(-1), (self.d,)) \
            if np.size(kp["k_length"]) in (1, self.d) else np.nan
        th[self.d] = float(np.asarray(kp["k_scale"]).reshape(-1)[0])
        th[self.d + 1] = float(np.asarray(noise).reshape(-1)[0])
        if self.kind == "Periodic":
            if kp.get("period") is None:
                raise ValueError("the Periodic kernel needs 'period' in the dict kernel_prior returns")
            th[self.d + 2] = float(np.asarray(kp["period"]).reshape(-1)[0])
        mean = None
        if self.has_mean_params:
            mean = np.asarray(self.m.mean_fn(self.X, mp), dtype=np.float64).squeeze()
        elif self.fixed_mean is not None:
            mean = self.fixed_mean
        return th, mean, sites

    def init_u(self):
        """init_to_median (gp.py:208)"""
        return np.concatenate([np.full(s.size, float(s.prior.inverse(s.prior.median()))) for s in self.sites])

    def _log_prior(self, u, sites, jacobian, want_grad=True):
        val, grad, o = 0.0, np.zeros(self.dim), 0
        for s0 in self.sites:
            pr = sites[s0.name].prior
            uu = u[o:o + s0.size]
            t = np.asarray(pr.transform(uu), dtype=np.float64)
            val += float(np.sum(pr.log_prob(t)))
            if want_grad:
                grad[o:o + s0.size] = np.asarray(pr.dlog_prob(t)) * np.asarray(pr.dtheta_du(uu))
            if jacobian:
                val += float(np.sum(pr.log_abs_jac(uu)))
                if want_grad:
                    grad[o:o + s0.size] += np.asarray(pr.dlog_abs_jac(uu))
            o += s0.size
        return val, grad

    def _lik(self, th, yres):
        """log p(y | theta, m) with d/dlog(theta) and alpha = K^-1 (y - m)"""
        if self._lik_fn is not None:
            return self._lik_fn(th, yres)
        val, g, alpha, info = self.m.ctx.mll(self.kind, self.X, yres, th, self.jitter, want_grad=True,
                                             want_alpha=self.has_mean_params)
        return val, g, alpha, info

    def __call__(self, u, jacobian):
        u = np.asarray(u, dtype=np.float64)
        th, mean, sites = self._run(u)
        self.n_evals += 1
        if not (np.all(np.isfinite(th)) and np.all(th[:self.d + 2] > 0)):
            return -np.inf, np.zeros(self.dim)
        yres = self.y0 if mean is None else self.y0 - mean
        val, g, alpha, info = self._lik(th, yres)
        if info != 0 or not np.isfinite(val):
            return -np.inf, np.zeros(self.dim)
        # site densities: analytic gradient, unless the programs are hierarchical (then differenced with the rest below)
        lp, grad = self._log_prior(u, sites, jacobian, want_grad=not self.hierarchical)
        # chain rule through the host programs by central differences
        h = self.FD_STEP
        dval_dth = g / th                                   # g is d/dlog(theta); unused entries of theta carry g = 0
        if self.has_mean_params and alpha is None:
            raise NotImplementedError("this likelihood does not return d value / d mean: no probabilistic mean function")
        for k in range(self.dim):
            e = np.zeros(self.dim)
            e[k] = h
            thp, mp_, sp = self._run(u + e)
            thm, mm_, sm_ = self._run(u - e)
            grad[k] += float(np.dot(dval_dth, (thp - thm) / (2 * h)))
            if self.has_mean_params:
                grad[k] += float(np.dot(alpha, (mp_ - mm_) / (2 * h)))
            if self.hierarchical:
                lpp, _ = self._log_prior(u + e, sp, jacobian, want_grad=False)
                lpm, _ = self._log_prior(u - e, sm_, jacobian, want_grad=False)
                grad[k] += (lpp - lpm) / (2 * h)
        return val + lp, grad

    def to_dict(self, U):
        """rows of unconstrained vectors -> dict of constrained site values, shapes as the program sampled them"""
        U = np.atleast_2d(U)
        out = {s.name: np.empty((U.shape[0],) + s.shape) for s in self.sites}
        for r, u in enumerate(U):
            o = 0
            for s in self.sites:
                out[s.name][r] = np.asarray(s.prior.transform(u[o:o + s.size])).reshape(s.shape)
                o += s.size
        return out


def make_log_joint(model, jitter=1e-6):
    if model.kernel_prior is not None or model.noise_prior is not None or \
            (model.mean_fn is not None and model.mean_fn_prior is not None):
        return ProgramLogJoint(model, jitter)
    return LogJoint(model, jitter)


# ---------------------------------------------------------------------------------------------- SVI
class SVIState:
    def __init__(self, losses, guide, loc, scale):
        self.losses, self.guide, self.loc, self.scale = losses, guide, loc, scale


def fit_vi_gp(model, rng_key, num_steps, step_size, progress_bar, **kwargs):
    """vigp.py:108-120: Adam(step_size, b1=0.5), AutoDelta (MAP in the constrained space, no Jacobian) or AutoNormal
    (mean-field normal over u, init scale 0.1, one reparameterised draw per step).  Returns (state, median dict)."""
    lj = make_log_joint(model, kwargs.get("jitter", 1e-6))
    rng = seed_from_key(rng_key)
    normal = model.guide_type == "normal"
    loc = lj.init_u()
    rho = np.full(lj.dim, math.log(0.1))            # log sigma
    params = np.concatenate([loc, rho]) if normal else loc.copy()
    m1, m2 = np.zeros_like(params), np.zeros_like(params)
    b1, b2, eps = 0.5, 0.999, 1e-8
    losses = []
    for t in range(1, int(num_steps) + 1):
        if normal:
            mu, r = params[:lj.dim], params[lj.dim:]
            e = rng.standard_normal(lj.dim)
            val, g = lj(mu + np.exp(r) * e, jacobian=True)
            elbo = val + r.sum() + 0.5 * lj.dim * (1 + math.log(2 * math.pi))
            grad = np.concatenate([g, g * e * np.exp(r) + 1.0])
        else:
            elbo, grad = lj(params, jacobian=False)
        losses.append(-elbo)
        if not np.isfinite(elbo):
            grad = np.zeros_like(params)
        m1 = b1 * m1 + (1 - b1) * (-grad)
        m2 = b2 * m2 + (1 - b2) * grad * grad
        params = params - step_size * (m1 / (1 - b1 ** t)) / (np.sqrt(m2 / (1 - b2 ** t)) + eps)
        if progress_bar and (t % max(1, num_steps // 10) == 0 or t == num_steps):
            print(f"svi step {t}/{num_steps}  loss {losses[-1]:.4f}")
    loc = params[:lj.dim]
    med = {k: (v[0] if v.ndim == 1 else v[0]) for k, v in lj.to_dict(loc).items()}   # guide median (vigp.py:125-127)
    return SVIState(np.array(losses), "normal" if normal else "delta", loc, np.exp(params[lj.dim:]) if normal else None), med


class SparseLogJoint(LogJoint):
    """LogJoint whose likelihood term is the VFE bound of viSparseGP.model (sparse_gp.py:62-114) at the current
    inducing inputs `self.Xu`; the gradient w.r.t. Xu of the last evaluation is left in `self.grad_Xu`."""

    def __init__(self, model, Xu, jitter=1e-6):
        super().__init__(model, jitter)
        self.Xu = np.array(Xu, dtype=np.float64, copy=True)
        if self.Xu.ndim == 1:
            self.Xu = self.Xu[:, None]
        self.grad_Xu = np.zeros_like(self.Xu)

    def _lik(self, th):
        val, g, gx, info = self.m.ctx.sparse_elbo(self.kind, self.Xu, self.X, self.y, th, self.jitter)
        self.grad_Xu = gx
        return val, g, info


def _sparse_program_log_joint(model, Xu0, jitter):
    """ProgramLogJoint over the VFE bound; carries `Xu` / `grad_Xu` like SparseLogJoint.  (The bound's derivative w.r.t.
    the mean vector is not computed on the device, so a probabilistic mean function is refused.)"""
    if model.mean_fn is not None and model.mean_fn_prior is not None:
        raise NotImplementedError("viSparseGP.fit with a probabilistic mean function is not implemented")
    lj = ProgramLogJoint(model, jitter, lik=None)
    lj.Xu = np.array(Xu0, dtype=np.float64, copy=True)
    if lj.Xu.ndim == 1:
        lj.Xu = lj.Xu[:, None]
    lj.grad_Xu = np.zeros_like(lj.Xu)

    def lik(th, yres):
        val, g, gx, info = model.ctx.sparse_elbo(lj.kind, lj.Xu, lj.X, yres, th, lj.jitter)
        lj.grad_Xu = gx
        return val, g, None, info
    lj._lik_fn = lik
    return lj


def fit_sparse_gp(model, rng_key, Xu0, num_steps, step_size, progress_bar, **kwargs):
    """sparse_gp.py:116-171: SVI with Adam(b1=0.5) over the hyper-parameters (delta or normal guide, as viGP) and over
    the inducing inputs Xu (a plain parameter, no prior).  Returns (state, dict with the guide median and 'Xu')."""
    if model.kernel_prior is not None or model.noise_prior is not None or \
            (model.mean_fn is not None and model.mean_fn_prior is not None):
        lj = _sparse_program_log_joint(model, Xu0, kwargs.get("jitter", 1e-6))
    else:
        lj = SparseLogJoint(model, Xu0, kwargs.get("jitter", 1e-6))
    rng = seed_from_key(rng_key)
    normal = model.guide_type == "normal"
    nu = lj.dim
    loc = lj.init_u()
    rho = np.full(nu, math.log(0.1))
    head = np.concatenate([loc, rho]) if normal else loc.copy()
    params = np.concatenate([head, lj.Xu.ravel()])
    nh = head.size
    m1, m2 = np.zeros_like(params), np.zeros_like(params)
    b1, b2, eps = 0.5, 0.999, 1e-8
    losses = []
    for t in range(1, int(num_steps) + 1):
        lj.Xu = params[nh:].reshape(lj.Xu.shape)
        if normal:
            mu, r = params[:nu], params[nu:nh]
            e = rng.standard_normal(nu)
            val, g = lj(mu + np.exp(r) * e, jacobian=True)
            elbo = val + r.sum() + 0.5 * nu * (1 + math.log(2 * math.pi))
            ghead = np.concatenate([g, g * e * np.exp(r) + 1.0])
        else:
            elbo, ghead = lj(params[:nu], jacobian=False)
        grad = np.concatenate([ghead, lj.grad_Xu.ravel()])
        losses.append(-elbo)
        if not np.isfinite(elbo):
            grad = np.zeros_like(params)
        m1 = b1 * m1 + (1 - b1) * (-grad)
        m2 = b2 * m2 + (1 - b2) * grad * grad
        params = params - step_size * (m1 / (1 - b1 ** t)) / (np.sqrt(m2 / (1 - b2 ** t)) + eps)
        if progress_bar and (t % max(1, num_steps // 10) == 0 or t == num_steps):
            print(f"svi step {t}/{num_steps}  loss {losses[-1]:.4f}")
    med = {k: v[0] for k, v in lj.to_dict(params[:nu]).items()}
    med["Xu"] = params[nh:].reshape(lj.Xu.shape)
    return SVIState(np.array(losses), "normal" if normal else "delta", params[:nu], np.exp(params[nu:nh]) if normal else None), med


# ---------------------------------------------------------------------------------------------- NUTS
class MCMCResult:
    """What ExactGP needs from numpyro's MCMC object: get_samples(group_by_chain)."""

    def __init__(self, samples_by_chain, stats):
        self._s, self.stats = samples_by_chain, stats

    def get_samples(self, group_by_chain=False):
        if group_by_chain:
            return self._s
        return {k: v.reshape((-1,) + v.shape[2:]) for k, v in self._s.items()}


def _leapfrog(lj, u, r, g, eps, minv):
    r = r + 0.5 * eps * g
    u = u + eps * minv * r
    lp, g = lj(u, jacobian=True)
    r = r + 0.5 * eps * g
    return u, r, lp, g


def _find_eps(lj, u, lp, g, rng, minv):
    eps = 1.0
    r = rng.standard_normal(u.size) / np.sqrt(minv)
    h0 = lp - 0.5 * np.dot(r, minv * r)
    _, r1, lp1, _ = _leapfrog(lj, u, r, g, eps, minv)
    h1 = lp1 - 0.5 * np.dot(r1, minv * r1)
    a = 1.0 if (np.isfinite(h1) and h1 - h0 > math.log(0.5)) else -1.0
    for _ in range(50):
        if not (np.isfinite(h1) and a * (h1 - h0) > -a * math.log(2)):
            if np.isfinite(h1) or a < 0:
                break
        eps *= 2.0 ** a
        _, r1, lp1, _ = _leapfrog(lj, u, r, g, eps, minv)
        h1 = lp1 - 0.5 * np.dot(r1, minv * r1)
    return eps


def _nuts_draw(lj, u0, lp0, g0, eps, rng, minv, max_depth=10):
    """one transition of the no-U-turn sampler with multinomial sampling along the trajectory"""
    r0 = rng.standard_normal(u0.size) / np.sqrt(minv)
    h0 = lp0 - 0.5 * np.dot(r0, minv * r0)
    um, rm, gm = u0.copy(), r0.copy(), g0.copy()
    up, rp, gp = u0.copy(), r0.copy(), g0.copy()
    u, lp, g = u0.copy(), lp0, g0.copy()
    logw = 0.0           # log of the total weight of the current tree (relative to h0)
    depth, alpha_sum, n_alpha, diverged = 0, 0.0, 0, False

    def build(u_, r_, g_, v, j):
        nonlocal alpha_sum, n_alpha, diverged
        if j == 0:
            u1, r1, lp1, g1 = _leapfrog(lj, u_, r_, g_, v * eps, minv)
            h1 = lp1 - 0.5 * np.dot(r1, minv * r1) if np.isfinite(lp1) else -np.inf
            dh = h1 - h0
            if not np.isfinite(dh):
                dh = -np.inf
            alpha_sum += min(1.0, math.exp(min(dh, 0.0))) if dh > -np.inf else 0.0
            n_alpha += 1
            ok = dh > -1000.0
            if not ok:
                diverged = True
            return u1, r1, g1, u1, r1, g1, u1, lp1, g1, dh, ok
        a = build(u_, r_, g_, v, j - 1)
        um_, rm_, gm_, up_, rp_, gp_, uc, lpc, gc, lw, ok = a
        if not ok:
            return a
        if v == -1:
            b = build(um_, rm_, gm_, v, j - 1)
            um_, rm_, gm_ = b[0], b[1], b[2]
        else:
            b = build(up_, rp_, gp_, v, j - 1)
            up_, rp_, gp_ = b[3], b[4], b[5]
        lw2, ok2 = b[9], b[10]
        lw_tot = np.logaddexp(lw, lw2)
        if ok2 and math.log(rng.uniform()) < lw2 - lw_tot:
            uc, lpc, gc = b[6], b[7], b[8]
        span = up_ - um_
        ok = ok2 and np.dot(span, minv * rm_) >= 0 and np.dot(span, minv * rp_) >= 0     # U-turn in the metric's velocity
        return um_, rm_, gm_, up_, rp_, gp_, uc, lpc, gc, lw_tot, ok

    while depth < max_depth:
        v = 1 if rng.uniform() < 0.5 else -1
        if v == -1:
            t = build(um, rm, gm, v, depth)
            um, rm, gm = t[0], t[1], t[2]
        else:
            t = build(up, rp, gp, v, depth)
            up, rp, gp = t[3], t[4], t[5]
        lw2, ok = t[9], t[10]
        if ok and math.log(rng.uniform()) < lw2 - logw:      # biased progressive sampling
            u, lp, g = t[6], t[7], t[8]
        logw = np.logaddexp(logw, lw2)
        span = up - um
        if not ok or np.dot(span, minv * rm) < 0 or np.dot(span, minv * rp) < 0:
            break
        depth += 1
    return u, lp, g, alpha_sum / max(n_alpha, 1), depth, diverged


def fit_exact_gp(model, rng_key, num_warmup, num_samples, num_chains, progress_bar, **kwargs):
    """gp.py:207-218."""
    return run_nuts(make_log_joint(model, kwargs.get("jitter", 1e-6)), rng_key, num_warmup, num_samples, num_chains, progress_bar)


def run_nuts(lj, rng_key, num_warmup, num_samples, num_chains, progress_bar):
    """NUTS over any log joint `lj(u, jacobian) -> (value, grad)` with `dim`, `init_u()`, `to_dict(U)`.  Dual-averaging
    step-size adaptation (target accept 0.8).  Warm-up windows as in Stan / NumPyro: draws of the middle of warm-up
    estimate a diagonal mass matrix, which is installed at 3/4 of warm-up; the step size is then re-initialised for the
    new metric and dual averaging restarts over the last quarter.  Chains run one after another ('sequential')."""
    root = seed_from_key(rng_key)
    chains, stats = [], []
    for c in range(int(num_chains)):
        rng = np.random.default_rng(root.integers(0, 2 ** 63))
        u = lj.init_u() + (0.0 if c == 0 else 0.1 * rng.standard_normal(lj.dim))
        lp, g = lj(u, jacobian=True)
        minv = np.ones(lj.dim)
        eps = _find_eps(lj, u, lp, g, rng, minv)
        mu, hbar, log_eps_bar, gamma, t0, kappa, delta = math.log(10 * eps), 0.0, 0.0, 0.05, 10.0, 0.75, 0.8
        draws, warm_buf, div, m = [], [], 0, 0
        metric_at = (3 * int(num_warmup)) // 4 if num_warmup >= 40 else -1
        for it in range(int(num_warmup) + int(num_samples)):
            u, lp, g, acc, depth, dv = _nuts_draw(lj, u, lp, g, eps, rng, minv)
            if it < num_warmup:
                m += 1
                hbar = (1 - 1 / (m + t0)) * hbar + (delta - acc) / (m + t0)
                log_eps = mu - math.sqrt(m) / gamma * hbar
                eta = m ** (-kappa)
                log_eps_bar = eta * log_eps + (1 - eta) * log_eps_bar
                eps = math.exp(log_eps)
                if num_warmup // 4 <= it < metric_at:
                    warm_buf.append(u.copy())
                if it == metric_at - 1 and len(warm_buf) >= 10:
                    var = np.var(np.array(warm_buf), axis=0)
                    minv = (len(warm_buf) * var + 1e-3 * 5) / (len(warm_buf) + 5)        # regularised, as Stan
                    eps = _find_eps(lj, u, lp, g, rng, minv)                              # step size for the new metric
                    mu, hbar, log_eps_bar, m = math.log(10 * eps), 0.0, 0.0, 0            # dual averaging restarts
                if it == num_warmup - 1:
                    eps = math.exp(log_eps_bar) if m > 0 else eps
            else:
                draws.append(u.copy())
                div += int(dv)
            if progress_bar and (it + 1) % max(1, (num_warmup + num_samples) // 10) == 0:
                print(f"chain {c} iter {it + 1}/{num_warmup + num_samples} step {eps:.3g} depth {depth} lp {lp:.3f}")
        chains.append(np.array(draws))
        stats.append({"step_size": eps, "divergences": div, "grad_evals": lj.n_evals})
    per_chain = [lj.to_dict(ch) for ch in chains]
    by_chain = {k: np.stack([pc[k] for pc in per_chain]) for k in per_chain[0]}
    return MCMCResult(by_chain, stats)
