"""
utils.py -- host helpers of the reference that the posterior path touches (gpax/utils/utils.py),
restated on NumPy.  Nothing here is numerical work.
"""
import numpy as np


_X64 = False


def enable_x64():
    """gpax/utils/utils.py:19-21.  The B200 path always computes in fp64; what the switch changes here is what it
    changes in the reference's random stream: `jax.random.normal` draws float64 variates from 64-bit words instead of
    float32 ones from 32-bit words, so y_sampled for a given key follows the reference in either mode."""
    global _X64
    _X64 = True


def x64_enabled() -> bool:
    return _X64


def get_keys(seed: int = 0):
    """gpax/utils/utils.py:24-30: `jax.random.split(jax.random.PRNGKey(seed))` -- the same two uint32[2] keys
    (prng.py restates JAX's threefry key derivation)."""
    from .prng import PRNGKey, split
    k = split(PRNGKey(seed), 2)
    return k[0], k[1]


def posterior_eps(rng_key, num_draws, n, P, dtype=np.float32, per_draw_keys=True):
    """Standard normals behind y_sampled, shape (num_draws, n, P), float64.

    An int seed or a JAX-style uint32[2] key reproduces the reference's stream: one sub-key per hyper-parameter draw
    (`jra.split(rng_key, num_samples)`, gp.py:391), `jax.random.normal(key, (n, P))` in the precision the reference
    runs in (float32 unless `enable_x64()` was called, as there).  `per_draw_keys=False` is the
    single-draw `_predict` (gp.py:279-293), which hands `rng_key` to the sampler as it is.  A numpy Generator (or
    None) is this package's own extension and draws from that generator."""
    if rng_key is None or isinstance(rng_key, np.random.Generator):
        return seed_from_key(rng_key).standard_normal((num_draws, n, P))
    from . import prng
    if per_draw_keys:
        return prng.mvn_eps(rng_key, num_draws, n, P, dtype)
    return prng.normal(prng.as_key(rng_key), (n, P), dtype).astype(np.float64)[None]


def seed_from_key(rng_key):
    """Accept an int, a JAX-style uint32[2] key, a numpy Generator, or None."""
    if rng_key is None:
        return np.random.default_rng(0)
    if isinstance(rng_key, np.random.Generator):
        return rng_key
    arr = np.asarray(rng_key)
    if arr.dtype == object:
        raise TypeError("rng_key must be an int or an integer array")
    return np.random.default_rng([int(v) & 0xFFFFFFFF for v in arr.reshape(-1)])


def split_in_batches(X_new, batch_size: int = 100, dim: int = 0):
    """gpax/utils/utils.py:33-51.  Same chunks as the reference for X_new.shape[dim] >= batch_size; the
    reference raises UnboundLocalError when there are fewer rows than batch_size --
    here that case returns the single short chunk."""
    if dim not in [0, 1]:
        raise NotImplementedError("'dim' must be equal to 0 or 1")
    n = X_new.shape[dim]
    num_batches = n // batch_size
    out = []
    for i in range(num_batches):
        out.append(X_new[i * batch_size:(i + 1) * batch_size] if dim == 0 else X_new[:, i * batch_size:(i + 1) * batch_size])
    start = num_batches * batch_size
    rest = X_new[start:] if dim == 0 else X_new[:, start:]
    if rest.shape[dim] > 0:
        out.append(rest)
    return out


def initialize_inducing_points(X, ratio=0.1, method="uniform", key=None):
    """gpax/utils/utils.py:171-212.  'uniform' uses proper integer indices (the reference casts them to
    int8 and wraps beyond 127 rows); 'random' draws without replacement from `key`."""
    if not 0 < ratio < 1:
        raise ValueError("The 'ratio' value must be between 0 and 1")
    X = np.asarray(X)
    n = X.shape[0]
    m = int(n * ratio)
    if method == "uniform":
        idx = np.linspace(0, n - 1, m).astype(np.int64)
        return X[idx]
    if method == "random":
        if key is None:
            raise ValueError("A random key must be provided for random selection")
        idx = seed_from_key(key).choice(n, size=m, replace=False)
        return X[idx]
    if method == "kmeans":
        from sklearn.cluster import KMeans
        return np.asarray(KMeans(n_clusters=m, random_state=0, n_init=10).fit(X).cluster_centers_)
    raise ValueError("Method must be 'uniform', 'random', or 'kmeans'")
