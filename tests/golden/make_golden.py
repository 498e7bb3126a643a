"""
make_golden.py -- generate tests/golden/*.npz by executing the REFERENCE'S OWN SOURCE FILES
(/root/reference/gpax, unmodified, read-only) in this container.

jax / jaxlib / numpyro / haiku / jaxopt / flax are not installable here (no network), so this
script installs stub modules for them before importing `gpax`:

  * `jax.numpy`            -> NumPy (fp64), with arrays returned as a tiny ndarray subclass that
                              offers the `.at[idx].add(v)` accessor sparse_gp.py:200 uses
  * `jax.scipy.linalg`     -> scipy.linalg (`cholesky`, `solve_triangular`)
  * `jax.jit`              -> identity decorator;  `jax.vmap` -> Python loop over axis 0
  * everything else (numpyro, jaxlib, haiku, ...) -> permissive dummies: those are only touched
    at import time (class definitions, type annotations), never on the posterior path.

What runs for real is the reference's Python: `gpax.kernels.RBFKernel/MaternKernel/PeriodicKernel`,
`gpax.ExactGP.get_mvn_posterior` (gp.py:253-277), `gpax.viGP.predict` (vigp.py:153-185),
`gpax.viSparseGP.get_mvn_posterior` (sparse_gp.py:173-223), `gpax.utils.split_in_batches`.
The arithmetic underneath is NumPy/LAPACK instead of XLA -- the same operations in the same order;
differences from a true JAX run are rounding-level (stated in DESIGN.md).

Run:  python tests/golden/make_golden.py        (needs /root/reference; not run on the GPU box)
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import scipy.linalg

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------- jax.numpy shim
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Upd:
            def add(self, v):
                out = np.array(arr, copy=True)
                out[idx] += v
                return out.view(JArr)

            def set(self, v):
                out = np.array(arr, copy=True)
                out[idx] = v
                return out.view(JArr)
        return _Upd()


class JArr(np.ndarray):
    @property
    def at(self):
        return _At(self)


def _wrap(fn):
    def f(*a, **k):
        r = fn(*a, **k)
        if isinstance(r, np.ndarray):
            return r.view(JArr)
        return r
    f.__name__ = getattr(fn, "__name__", "f")
    return f


def _make_jnp():
    m = types.ModuleType("jax.numpy")
    for name in dir(np):
        if name.startswith("_"):
            continue
        obj = getattr(np, name)
        if isinstance(obj, (np.ufunc,)) or (callable(obj) and not isinstance(obj, type)):
            setattr(m, name, _wrap(obj))
        else:
            setattr(m, name, obj)
    m.ndarray = np.ndarray
    m.array = _wrap(np.array)
    lin = types.ModuleType("jax.numpy.linalg")
    for name in ("inv", "cholesky", "solve", "det", "norm", "slogdet"):
        setattr(lin, name, _wrap(getattr(np.linalg, name)))
    m.linalg = lin
    return m, lin


class _Dummy:
    """Permissive stand-in: any attribute, call, subscript or subclassing works."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __getitem__(self, k):
        return _Dummy()

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return _Dummy()


_STUB_ROOTS = ("jax", "jaxlib", "numpyro", "haiku", "jaxopt", "flax", "optax")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def install_shim():
    jnp, lin = _make_jnp()
    jax = _StubModule("jax")
    jax.__path__ = []
    jax.numpy = jnp
    jax.jit = lambda f=None, **k: f if f is not None else (lambda g: g)

    def vmap(f, in_axes=0, out_axes=0):
        def g(*args):
            n = len(args[0])
            outs = [f(*[a[i] for a in args]) for i in range(n)]
            return np.stack(outs)
        return g
    jax.vmap = vmap
    jsl = types.ModuleType("jax.scipy.linalg")
    jsl.cholesky = _wrap(lambda a, lower=False, **k: scipy.linalg.cholesky(a, lower=lower))
    jsl.solve_triangular = _wrap(lambda a, b, lower=False, trans=0, **k:
                                 scipy.linalg.solve_triangular(a, b, lower=lower, trans=trans))
    jsc = _StubModule("jax.scipy")
    jsc.__path__ = []
    jsc.linalg = jsl
    jax.scipy = jsc
    sys.modules.update({"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": lin,
                        "jax.scipy": jsc, "jax.scipy.linalg": jsl})
    sys.meta_path.insert(0, _StubFinder())


# ----------------------------------------------------------------------------- cases
def dummy_data(rng, n=8):
    """Shape of the reference's own fixture, tests/test_gp.py:15-22 (there unseeded)."""
    X = np.linspace(1, 2, n) + 0.1 * rng.standard_normal(n)
    y = 10 * X ** 2
    return X, y


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; golden vectors are generated in the build container only")
    install_shim()
    sys.path.insert(0, REF)
    import gpax  # the reference package, unmodified

    rng = np.random.default_rng(20260924)
    out = {}

    # --- Gram matrices: scalar and ARD lengthscales, d in {1,2,3}, X==Z-shape and X!=Z-shape
    kcases = []
    for kname, fn in (("RBF", gpax.kernels.RBFKernel), ("Matern", gpax.kernels.MaternKernel),
                      ("Periodic", gpax.kernels.PeriodicKernel)):
        for d in (1, 2, 3):
            for n, m in ((5, 5), (7, 4), (33, 33)):
                X = rng.uniform(-1, 2, (n, d))
                Z = rng.uniform(-1, 2, (m, d))
                ell = rng.uniform(0.3, 1.5, d) if d > 1 else np.array([0.7])
                params = {"k_length": ell.view(JArr), "k_scale": np.float64(1.3),
                          "period": np.float64(0.9)}
                noise = 0.05
                K = np.asarray(fn(X.view(JArr), Z.view(JArr), params, noise, jitter=1e-6))
                tag = f"gram_{kname}_d{d}_{n}x{m}"
                out[tag + "_X"], out[tag + "_Z"], out[tag + "_ell"], out[tag + "_K"] = X, Z, ell, K
                kcases.append(tag)
    # same-array train/train Gram (diagonal exactly k(x,x)+noise+jitter)
    X = rng.uniform(0, 1, (16, 2))
    params = {"k_length": np.array([0.4, 0.6]).view(JArr), "k_scale": np.float64(2.0), "period": np.float64(1.0)}
    for kname, fn in (("RBF", gpax.kernels.RBFKernel), ("Matern", gpax.kernels.MaternKernel),
                      ("Periodic", gpax.kernels.PeriodicKernel)):
        out[f"gramself_{kname}_K"] = np.asarray(fn(X.view(JArr), X.view(JArr), params, 0.1, jitter=1e-6))
    out["gramself_X"] = X

    # --- ExactGP.get_mvn_posterior, the reference test's own setting (tests/test_gp.py:139-152)
    Xtr, ytr = dummy_data(rng, 8)
    Xte = np.linspace(1, 2, 20)
    for kname in ("RBF", "Matern", "Periodic"):
        m = gpax.ExactGP(1, kernel=kname)
        m.X_train = Xtr[:, None].view(JArr)
        m.y_train = ytr.view(JArr)
        params = {"k_length": np.array([1.0]).view(JArr), "k_scale": np.float64(1.0),
                  "noise": np.float64(0.1), "period": np.float64(1.0)}
        for noiseless in (False, True):
            mean, cov = m.get_mvn_posterior(Xte[:, None].view(JArr), params, noiseless)
            out[f"exact8_{kname}_nl{int(noiseless)}_mean"] = np.asarray(mean)
            out[f"exact8_{kname}_nl{int(noiseless)}_cov"] = np.asarray(cov)
        mean, cov = m.get_mvn_posterior(Xte[:, None].view(JArr), params, False, jitter=1e-5)
        out[f"exact8_{kname}_jit1e-5_mean"], out[f"exact8_{kname}_jit1e-5_cov"] = np.asarray(mean), np.asarray(cov)
    out["exact8_Xtr"], out["exact8_ytr"], out["exact8_Xte"] = Xtr, ytr, Xte

    # --- a larger, well-conditioned 2-D / 3-D problem per kernel (ARD), plus viGP.predict
    for kname, d, N, P in (("RBF", 3, 300, 40), ("Matern", 2, 384, 50), ("Periodic", 1, 200, 30)):
        Xtr = rng.uniform(0, 1, (N, d))
        ytr = np.sin(4 * Xtr[:, 0]) * np.cos(3 * Xtr[:, -1]) + 0.1 * rng.standard_normal(N)
        Xte = rng.uniform(0, 1, (P, d))
        ell = np.full(d, 0.3) * (1 + 0.2 * np.arange(d))
        params = {"k_length": ell.view(JArr), "k_scale": np.float64(1.2),
                  "noise": np.float64(0.1), "period": np.float64(0.8)}
        m = gpax.ExactGP(d, kernel=kname)
        m.X_train, m.y_train = Xtr.view(JArr), ytr.view(JArr)
        mean, cov = m.get_mvn_posterior(Xte.view(JArr), params, False)
        tag = f"exact_{kname}_N{N}"
        out[tag + "_Xtr"], out[tag + "_ytr"], out[tag + "_Xte"], out[tag + "_ell"] = Xtr, ytr, Xte, ell
        out[tag + "_mean"], out[tag + "_cov"] = np.asarray(mean), np.asarray(cov)
        v = gpax.viGP(d, kernel=kname)
        v.X_train, v.y_train = Xtr.view(JArr), ytr.view(JArr)
        vm, vv = v.predict(None, Xte.view(JArr), samples=params, noiseless=True)
        out[tag + "_vimean"], out[tag + "_vivar"] = np.asarray(vm), np.asarray(vv)

    # --- mean function with parameters (gp.py:262-265, 274-276)
    Xtr, ytr = dummy_data(rng, 12)
    Xte = np.linspace(1, 2, 9)
    mfn = lambda x, p: p["a"] * x[:, 0] ** 2 + p["b"]           # noqa: E731
    m = gpax.ExactGP(1, kernel="RBF", mean_fn=mfn, mean_fn_prior=lambda: None)
    m.X_train, m.y_train = Xtr[:, None].view(JArr), ytr.view(JArr)
    params = {"k_length": np.array([0.5]).view(JArr), "k_scale": np.float64(1.0), "noise": np.float64(0.05),
              "a": np.float64(9.0), "b": np.float64(0.5)}
    mean, cov = m.get_mvn_posterior(Xte[:, None].view(JArr), params, False)
    out["meanfn_Xtr"], out["meanfn_ytr"], out["meanfn_Xte"] = Xtr, ytr, Xte
    out["meanfn_mean"], out["meanfn_cov"] = np.asarray(mean), np.asarray(cov)

    # --- viSparseGP.get_mvn_posterior (tests/test_sparsegp.py:47-64 setting: Xu = X[::2]) + a 2-D case
    for tag, N, d, kname in (("sparse50", 50, 1, "RBF"), ("sparse400", 400, 2, "Matern")):
        Xtr = rng.uniform(0, 1, (N, d))
        ytr = np.sin(6 * Xtr[:, 0]) + 0.1 * rng.standard_normal(N)
        Xu = Xtr[::2] if N == 50 else Xtr[rng.choice(N, 40, replace=False)]
        Xte = rng.uniform(0, 1, (25, d))
        ell = np.full(d, 0.4)
        params = {"k_length": ell.view(JArr), "k_scale": np.float64(1.0), "noise": np.float64(0.1)}
        m = gpax.viSparseGP(d, kernel=kname)
        m.X_train, m.y_train, m.Xu = Xtr.view(JArr), ytr.view(JArr), Xu.view(JArr)
        for noiseless in (False, True):
            mean, cov = m.get_mvn_posterior(Xte.view(JArr), params, noiseless, jitter=1e-5)
            out[f"{tag}_nl{int(noiseless)}_mean"] = np.asarray(mean)
            out[f"{tag}_nl{int(noiseless)}_cov"] = np.asarray(cov)
        out[tag + "_Xtr"], out[tag + "_ytr"], out[tag + "_Xu"], out[tag + "_Xte"] = Xtr, ytr, Xu, Xte

    # --- split_in_batches (utils.py:33-51): lengths for a few sizes
    A = np.arange(23.0)[:, None]
    for bs in (2, 3, 8, 23):
        parts = gpax.utils.split_in_batches(A.view(JArr), bs)
        out[f"split23_bs{bs}_lens"] = np.array([len(p) for p in parts])

    out["gram_cases"] = np.array(kcases)
    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(OUT, "reference_vectors.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
