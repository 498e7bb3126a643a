"""
inference.py -- hyper-parameter inference for `.fit()` (the row after the
posterior path).  Placeholder until the marginal-likelihood value+gradient kernel lands.
"""


def _todo(name):
    raise NotImplementedError(
        f"{name}: the fit path (NUTS / SVI over the marginal likelihood, gpax/models/gp.py:166-220, "
        "vigp.py:77-123) is the next row of the scope table; pass `samples=` / `params` to the predict "
        "path, which is what this build accelerates.")


def fit_exact_gp(model, rng_key, num_warmup, num_samples, num_chains, progress_bar, **kwargs):
    _todo("ExactGP.fit")


def fit_vi_gp(model, rng_key, num_steps, step_size, progress_bar, **kwargs):
    _todo("viGP.fit")


def fit_sparse_gp(model, rng_key, Xu0, num_steps, step_size, progress_bar, **kwargs):
    _todo("viSparseGP.fit")
