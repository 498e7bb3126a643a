"""torchrun worker: block-cyclic posterior and sharded sparse posterior on WORLD_SIZE GPUs vs the oracle.
Launched by tests/test_gpu_distributed.py (and usable by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_worker.py)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from gpax_b200.distributed import BlockCyclicGP, GpuOps, sharded_sparse_posterior  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = GpuOps(device=local)
    rng = np.random.default_rng(5)
    # ---- block-cyclic Cholesky posterior
    N, P, d, nb = 1500, 64, 3, 256
    X = rng.uniform(0, 1, (N, d))
    y = np.sin(4 * X[:, 0]) + X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(N)
    Xn = rng.uniform(0, 1, (P, d))
    theta = np.array([0.3, 0.35, 0.4, 1.2, 0.1, 1.0])
    gp = BlockCyclicGP(ops, N, nb)
    mean, var, info = gp.posterior("RBF", ops.from_numpy(X), ops.from_numpy(y), ops.from_numpy(Xn), theta)
    ref_mean, ref_cov = oracle.exact_posterior(X, y, Xn, {"k_length": theta[:3], "k_scale": 1.2, "noise": 0.1}, "RBF")
    assert info == 0
    np.testing.assert_allclose(mean, ref_mean, rtol=1e-9, atol=1e-9 * np.abs(ref_mean).max())
    np.testing.assert_allclose(var, np.diag(ref_cov), rtol=1e-9, atol=1e-9 * np.abs(ref_cov).max())
    if world > 1:
        assert gp.bytes_broadcast == sum((N - k * nb) * nb * 8 for k in range(gp.nblk))
    # ---- sharded sparse posterior
    N, M, P, d = 3000, 200, 50, 2
    X = rng.uniform(0, 1, (N, d))
    y = np.sin(5 * X[:, 0]) + 0.1 * rng.standard_normal(N)
    Xu = X[rng.choice(N, M, replace=False)]
    Xn = rng.uniform(0, 1, (P, d))
    theta = np.array([0.4, 0.4, 1.0, 0.1, 1.0])
    lo, hi = rank * N // world, (rank + 1) * N // world
    out = sharded_sparse_posterior(ops, "Matern", ops.from_numpy(Xu), ops.from_numpy(X[lo:hi]), ops.from_numpy(y[lo:hi]),
                                   ops.from_numpy(Xn), theta, jitter=1e-5, want_cov=True)
    ref_mean, ref_cov = oracle.sparse_posterior(X, y, Xu, Xn, {"k_length": theta[:2], "k_scale": 1.0, "noise": 0.1}, "Matern",
                                                jitter=1e-5)
    assert out["info"] == 0
    np.testing.assert_allclose(out["mean"], ref_mean, rtol=1e-6, atol=1e-6 * np.abs(ref_mean).max())
    np.testing.assert_allclose(out["cov"], ref_cov, rtol=1e-6, atol=1e-6 * np.abs(ref_cov).max())
    dist.barrier()
    if rank == 0:
        print(f"DIST_GPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
