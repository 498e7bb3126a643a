"""bench_multigpu.py -- the two configurations of BASELINE.json that need a collective, timed under torchrun:
   c4: ExactGP RBF N=32768 d=3 fp64, block-cyclic Cholesky across the ranks (NCCL panel broadcast)
   c5: viSparseGP N=262144 M=4096, Kuf build + Nystrom solve, training set sharded across the ranks (all-reduce)
 usage: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port 29520 tools/bench_multigpu.py c4|c5 [N] [nb]
 prints one JSON line on rank 0 (wall-clock bracketed by barriers + device synchronize, max over ranks)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpax_b200.distributed import BlockCyclicGP, GpuOps, sharded_sparse_posterior  # noqa: E402


def tmax(x, dev):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    which = sys.argv[1]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = GpuOps(device=local)
    if which == "c4":
        N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
        nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
        P, d = 1024, 3
        rng = np.random.default_rng(5)
        X = rng.uniform(0, 1, (N, d))
        y = np.sin(3 * X[:, 0]) + X[:, 1] * X[:, 2] + 0.1 * rng.standard_normal(N)
        Xn = rng.uniform(0, 1, (P, d))
        theta = np.array([0.3, 0.3, 0.3, 1.0, 0.1, 1.0])
        dX, dy, dXn = ops.from_numpy(X), ops.from_numpy(y), ops.from_numpy(Xn)
        gp = BlockCyclicGP(ops, N, nb)
        res = {}
        for rep in range(2):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            gp.build("RBF", dX, theta, 1e-6)
            torch.cuda.synchronize(); dist.barrier()
            t1 = time.perf_counter()
            info = gp.factor()
            torch.cuda.synchronize(); dist.barrier()
            t2 = time.perf_counter()
            mean, var = gp.solve_mean_var("RBF", dX, dy, dXn, theta, False, 1e-6)
            torch.cuda.synchronize(); dist.barrier()
            t3 = time.perf_counter()
            res = {"gram_ms": tmax(t1 - t0, dev) * 1e3, "potrf_ms": tmax(t2 - t1, dev) * 1e3, "solve_ms": tmax(t3 - t2, dev) * 1e3,
                   "total_ms": tmax(t3 - t0, dev) * 1e3, "info": info}
        if rank == 0:
            res.update({"config": f"c4 ExactGP RBF N={N} d=3 P={P} block-column-cyclic nb={nb}", "n_gpus": world,
                        "potrf_tflops_N3_3": N ** 3 / 3 / (res["potrf_ms"] / 1e3) / 1e12,
                        "posteriors_per_s": 1e3 / res["total_ms"], "bytes_broadcast_per_rank": gp.bytes_broadcast // 2,
                        "mean_finite": bool(np.isfinite(mean).all()), "var_min": float(var.min())})
            print(json.dumps(res))
    else:
        N = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
        M, P, d = 4096, 4096, 2
        rng = np.random.default_rng(6)
        lo, hi = rank * N // world, (rank + 1) * N // world
        Xfull = rng.uniform(0, 1, (N, d))
        yfull = np.sin(5 * Xfull[:, 0]) * np.cos(4 * Xfull[:, 1]) + 0.05 * rng.standard_normal(N)
        Xu = Xfull[rng.choice(N, M, replace=False)]
        Xn = rng.uniform(0, 1, (P, d))
        theta = np.array([0.2, 0.2, 1.0, 0.05, 1.0])
        dXu, dX, dy, dXn = ops.from_numpy(Xu), ops.from_numpy(Xfull[lo:hi]), ops.from_numpy(yfull[lo:hi]), ops.from_numpy(Xn)
        res = {}
        for rep in range(2):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sharded_sparse_posterior(ops, "Matern", dXu, dX, dy, dXn, theta, jitter=1e-5)
            torch.cuda.synchronize(); dist.barrier()
            res = {"total_ms": tmax(time.perf_counter() - t0, dev) * 1e3, "info": out["info"]}
        if rank == 0:
            flops = 2.0 * M * M * N + 2 * M ** 3 / 3 + 2.0 * M * M * (P + 1)
            res.update({"config": f"c5 viSparseGP Matern N={N} M={M} P={P} d=2, N sharded", "n_gpus": world,
                        "tflops": flops / (res["total_ms"] / 1e3) / 1e12, "posteriors_per_s": 1e3 / res["total_ms"],
                        "allreduce_bytes": 8 * (M * M + M), "mean_finite": bool(np.isfinite(out["mean"]).all()),
                        "var_min": float(out["var"].min())})
            print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
