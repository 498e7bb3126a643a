"""
vigp.py -- `viGP` with the reference's surface (gpax/models/vigp.py:61-75 constructor, 125-127
get_samples, 129-151 predict_in_batches, 153-185 predict).  predict() is one b2gp_posterior call with the
diagonal-variance epilogue: the P x P covariance the reference builds and throws away (vigp.py:184-185)
is never formed.
"""
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from .gp import ExactGP


class viGP(ExactGP):
    """Variational-inference GP: a single (MAP / guide-median) theta instead of HMC draws."""

    def __init__(self, input_dim: int, kernel, mean_fn: Optional[Callable] = None,
                 kernel_prior: Optional[Callable] = None, mean_fn_prior: Optional[Callable] = None,
                 noise_prior: Optional[Callable] = None, noise_prior_dist=None, lengthscale_prior_dist=None,
                 guide: str = "delta", ctx=None) -> None:
        super().__init__(input_dim, kernel, mean_fn, kernel_prior, mean_fn_prior, noise_prior,
                         noise_prior_dist, lengthscale_prior_dist, ctx=ctx)
        self.guide_type = "normal" if guide == "normal" else "delta"   # vigp.py:74
        self.svi = None
        self.kernel_params = None

    def fit(self, rng_key, X, y, num_steps: int = 1000, step_size: float = 5e-3, progress_bar: bool = True,
            print_summary: bool = True, device=None, **kwargs: float) -> None:
        """vigp.py:77-123: SVI with Adam(b1=0.5) and an AutoDelta / AutoNormal guide in the reference."""
        from .inference import fit_vi_gp
        X, y = self._set_data(X, y)
        self.X_train, self.y_train = X, y
        self.svi, self.kernel_params = fit_vi_gp(self, rng_key, num_steps, step_size, progress_bar, **kwargs)
        if print_summary:
            self._print_summary()

    def get_samples(self) -> Dict[str, np.ndarray]:
        """vigp.py:125-127: the guide's median."""
        if self.kernel_params is None:
            raise RuntimeError("no variational parameters: call fit() first or pass `samples=` to predict()")
        return self.kernel_params

    def predict(self, rng_key, X_new, samples: Optional[Dict[str, np.ndarray]] = None, noiseless: bool = False,
                device=None, **kwargs: float) -> Tuple[np.ndarray, np.ndarray]:
        """vigp.py:153-185: (mean [P], diag of the posterior covariance [P])."""
        X_new = self._set_data(X_new)
        if samples is None:
            samples = self.get_samples()
        dt = self._out_dtype(X_new)
        if self._fused is None:
            mean, cov = self._posterior_callable(X_new, samples, noiseless, **kwargs)
            return mean.astype(dt, copy=False), np.diagonal(cov).astype(dt, copy=False)
        out = self._posterior_batched(X_new, samples, False, noiseless, ("mean", "var"), **kwargs)
        return out["mean"][0].astype(dt, copy=False), out["var"][0].astype(dt, copy=False)

    def predict_in_batches(self, rng_key, X_new, batch_size: int = 100, samples=None, predict_fn=None,
                           noiseless: bool = False, device=None, **kwargs: float) -> Tuple[np.ndarray, np.ndarray]:
        """vigp.py:129-151.  The reference chunks X_new to bound the P x P covariance it forms per chunk and re-inverts
        k_XX for every chunk.  Here (mean, var) of a test point do not depend on which other points share its chunk, the
        factor is kept between chunks, and no P x P matrix exists, so chunks smaller than INTERNAL_BATCH rows are merged
        before they go to the device: same outputs, the triangular solve runs on machine-filling panels instead of
        `batch_size`-row slivers (default 100).  A user `predict_fn` keeps the caller's chunking."""
        if predict_fn is None:
            batch_size = max(int(batch_size), self.INTERNAL_BATCH)
            if np.asarray(X_new).shape[0] <= batch_size:        # split_in_batches would drop a short single chunk
                return self.predict(rng_key, X_new, samples, noiseless, **kwargs)
            predict_fn = lambda xi: self.predict(rng_key, xi, samples, noiseless, **kwargs)   # noqa: E731
        y_pred, y_var = self._predict_in_batches(rng_key, X_new, batch_size, 0, samples, predict_fn=predict_fn,
                                                 noiseless=noiseless, device=device, **kwargs)
        return np.concatenate(y_pred, 0), np.concatenate(y_var, 0)

    INTERNAL_BATCH = 8192

    def _print_summary(self) -> None:
        print("\nInferred GP parameters")
        for k, vals in self.get_samples().items():
            print(k, " " * (15 - len(k)), np.around(np.asarray(vals), 4))
