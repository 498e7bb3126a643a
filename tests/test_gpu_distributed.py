"""GPU: gpax_b200/distributed.py through the real C-ABI ops.  World size 1 runs in-process (no collective);
the 2-GPU case launches tests/dist_gpu_worker.py under torchrun when two devices are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_block_cyclic_single_gpu():
    from gpax_b200.distributed import BlockCyclicGP, GpuOps
    ops = GpuOps(device=0)
    rng = np.random.default_rng(3)
    N, P, d, nb = 1100, 40, 2, 256        # ragged last block column (1100 = 4*256 + 76)
    X = rng.uniform(0, 1, (N, d))
    y = np.sin(4 * X[:, 0]) * np.cos(3 * X[:, 1]) + 0.1 * rng.standard_normal(N)
    Xn = rng.uniform(0, 1, (P, d))
    theta = np.array([0.3, 0.4, 1.0, 0.1, 1.0])
    for kind in ("RBF", "Matern"):
        gp = BlockCyclicGP(ops, N, nb)
        mean, var, info = gp.posterior(kind, ops.from_numpy(X), ops.from_numpy(y), ops.from_numpy(Xn), theta)
        ref_mean, ref_cov = oracle.exact_posterior(X, y, Xn, {"k_length": theta[:2], "k_scale": 1.0, "noise": 0.1}, kind)
        assert info == 0
        np.testing.assert_allclose(mean, ref_mean, rtol=1e-9, atol=1e-9 * np.abs(ref_mean).max())
        np.testing.assert_allclose(var, np.diag(ref_cov), rtol=1e-9, atol=1e-9 * np.abs(ref_cov).max())
    # same answer as the single-call path
    import gpax_b200
    one = gpax_b200.default_context().posterior("Matern", X, y, Xn, theta[None], want=("mean", "var"))
    np.testing.assert_allclose(mean, one["mean"][0], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(var, one["var"][0], rtol=1e-10, atol=1e-12)


def test_sparse_partial_finish_single_gpu(golden):
    from gpax_b200.distributed import GpuOps, sharded_sparse_posterior
    ops = GpuOps(device=0)
    tag = "sparse400"
    Xtr, ytr, Xu, Xte = (golden[tag + s] for s in ("_Xtr", "_ytr", "_Xu", "_Xte"))
    theta = np.array([0.4, 0.4, 1.0, 0.1, 1.0])
    out = sharded_sparse_posterior(ops, "Matern", ops.from_numpy(Xu), ops.from_numpy(Xtr), ops.from_numpy(ytr),
                                   ops.from_numpy(Xte), theta, jitter=1e-5, want_cov=True)
    assert out["info"] == 0
    ref_mean, ref_cov = golden[f"{tag}_nl0_mean"], golden[f"{tag}_nl0_cov"]
    np.testing.assert_allclose(out["mean"], ref_mean, rtol=1e-6, atol=1e-6 * np.abs(ref_mean).max())
    np.testing.assert_allclose(out["cov"], ref_cov, rtol=1e-6, atol=1e-6 * np.abs(ref_cov).max())


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpus_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DIST_GPU_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
