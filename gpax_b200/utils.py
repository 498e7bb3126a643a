"""
utils.py -- host helpers of the reference that the posterior path touches (gpax/utils/utils.py),
restated on NumPy.  Nothing here is numerical work.
"""
import numpy as np


def enable_x64():
    """gpax/utils/utils.py:19-21.  The B200 path always computes in fp64; kept for drop-in use."""
    return None


def get_keys(seed: int = 0):
    """gpax/utils/utils.py:24-30: two keys for fit / predict.  Keys here are uint32[2] arrays used only
    as seeds (JAX's threefry stream is not reproduced)."""
    ss = np.random.SeedSequence(seed)
    a, b = ss.spawn(2)
    return a.generate_state(2).astype(np.uint32), b.generate_state(2).astype(np.uint32)


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
