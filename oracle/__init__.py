"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU (NumPy/SciPy, fp64) restatement of the reference's exact-GP posterior path.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package, and only as the checker or the timed CPU baseline -- never as
the product path.  `gpax_b200` never imports it.

Pinning status: the reference (gpax v0.1.9) needs jax/jaxlib/numpyro, none of which is
installed in the build image, and its own test-suite holds no golden value on this path
(SURVEY.md section 8c).  The oracle is therefore pinned against the reference's *own source
files executed over a NumPy shim of `jax.numpy`* (`tests/golden/make_golden.py`, which imports
/root/reference/gpax with stub `jax`/`numpyro` modules and runs the unmodified
`RBFKernel/MaternKernel/PeriodicKernel`, `ExactGP.get_mvn_posterior`,
`viGP.predict` and `viSparseGP.get_mvn_posterior`); the resulting vectors are committed under
`tests/golden/` and `tests/test_oracle_golden.py` checks the oracle against them.
It is NOT pinned against a real JAX/XLA execution ("parity unpinned" in that strict sense).
"""
from .gp_oracle import (  # noqa: F401
    square_scaled_distance, rbf_kernel, matern_kernel, periodic_kernel, get_kernel,
    exact_posterior, exact_posterior_chol, vi_predict, predict_draws,
    sparse_posterior, split_in_batches,
)
