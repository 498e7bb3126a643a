"""CPU: host-side logic of the section-8f shells that needs no GPU -- acquisition penalties (gpax/acquisition/penalties.py),
the task-covariance index kernel (gpax/kernels/mtkernels.py:19-58), argument checks."""
import numpy as np
import pytest

from gpax_b200 import acquisition as acq
from gpax_b200 import dist, mtkernels


def test_penalties_match_the_reference_formulas():
    X = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0], [3.0, 4.0]])
    recent = np.array([[1.0, 0.0], [3.0, 4.0]])
    d = acq.compute_penalty(X, recent, "delta")
    assert np.array_equal(np.isinf(d), [False, True, False, True]) and (d[[0, 2]] == 0).all()
    p = acq.compute_penalty(X, recent, "inverse_distance", penalty_factor=2.0)
    ts = np.arange(3, 1, -1)                       # penalties.py: timestamps [len+1 .. 2]
    ref = [2.0 * np.sum(1 / (np.linalg.norm(recent - x, axis=1) + 1) / ts) for x in X]
    np.testing.assert_allclose(p, ref, rtol=1e-15)
    one = acq.compute_penalty(X, recent[:1], "inverse_distance")
    np.testing.assert_allclose(one, 1 / (np.linalg.norm(recent[:1] - X, axis=1) + 1), rtol=1e-15)
    with pytest.raises(NotImplementedError):
        acq.compute_penalty(X, recent, "nope")


def test_index_kernel_and_grid_defaults():
    W = np.array([[1.0, 0.5], [0.2, -0.3], [0.0, 1.0]])
    v = np.array([0.1, 0.2, 0.3])
    B = W @ W.T + np.diag(v)
    k = mtkernels.index_kernel([0, 2, 2], [1, 0], {"W": W, "v": v})
    np.testing.assert_array_equal(k, B[np.ix_([0, 2, 2], [1, 0])])
    with pytest.raises(NotImplementedError):
        mtkernels.MultitaskKernel(lambda *a, **k: None)
    assert dist.default_grid(8) == (2, 4)
