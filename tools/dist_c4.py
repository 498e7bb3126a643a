"""dist_c4.py -- config 4 (ExactGP RBF N=32768 d=3 P=1024) through the in-library block-cyclic path on all visible GPUs.
 usage: python tools/dist_c4.py WORLD [nb] [PRxPC] [N]      (spawns WORLD ranks; B2GP_DIST_PROF=1 prints the phase profile)"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    from gpax_b200 import dist
    nb, grid, N = int(os.environ["C4_NB"]), os.environ["C4_GRID"], int(os.environ["C4_N"])
    pr, pc = (int(v) for v in grid.split("x"))
    dc = dist.DistContext(grid=(pr, pc))
    d, P = 3, 1024
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 1, (N, d))
    y = np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) + X[:, 2] + 0.1 * rng.standard_normal(N)
    Xn = rng.uniform(0, 1, (P, d))
    theta = np.array([0.3, 0.3, 0.3, 1.0, 0.1, 1.0])
    for k, v in (kv.split("=") for kv in os.environ.get("C4_OPTS", "").split() if kv):
        dc.ctx.set_option(k, int(v))
    res = None
    times = []
    for it in range(4):
        res = dc.posterior("RBF", X, y, Xn, theta, nb=nb)
        times.append((res["timing"]["total_ms"], res["timing"]["potrf_ms"]))
    if dc.rank == 0:
        print(json.dumps({"world": dc.world, "grid": grid, "nb": nb, "N": N, "total_ms": [round(t[0], 2) for t in times],
                          "factorisation_ms": [round(t[1], 2) for t in times], "info": res["info"],
                          "checksum": float(np.abs(res["mean"]).sum())}), flush=True)
    dc.close()


def main():
    world = int(sys.argv[1])
    nb = sys.argv[2] if len(sys.argv) > 2 else "512"
    from gpax_b200 import dist
    grid = sys.argv[3] if len(sys.argv) > 3 else "%dx%d" % dist.default_grid(world)
    N = sys.argv[4] if len(sys.argv) > 4 else "32768"
    port = 29700 + os.getpid() % 200
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   C4_NB=nb, C4_GRID=grid, C4_N=N, C4_WORKER="1")
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__)], env=env))
    rc = 0
    for p in procs:
        rc |= p.wait()
    sys.exit(rc)


if __name__ == "__main__":
    if os.environ.get("C4_WORKER"):
        worker()
    else:
        main()
