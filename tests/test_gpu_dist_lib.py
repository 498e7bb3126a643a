"""GPU: the in-library multi-GPU posterior (b2gp_dist_posterior, gpax_b200/csrc/dist.cuh) against the single-GPU posterior.
A 1 x 1 "grid" runs the whole block-cyclic machinery (tile lists, row maps, look-ahead order) on one GPU; with two or more
GPUs visible the same problem runs as one process per GPU over NCCL on 1 x 2, 2 x 1 (and 2 x 2 / 2 x 4) grids."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT, assert_close
from dist_lib_worker import problem

pytestmark = pytest.mark.gpu


def n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return len([l for l in out.splitlines() if l.startswith("GPU ")])
    except Exception:  # noqa: BLE001
        return 0


def run_grid(pr, pc, N, P, nb, kernel):
    world = pr * pc
    port = 29600 + (os.getpid() + 7 * pr + 13 * pc + N) % 300
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "res")
        procs = []
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_lib_worker.py"), str(pr), str(pc), str(N),
                                           str(P), str(nb), kernel, out], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        logs = [p.communicate(timeout=600)[0] for p in procs]
        for p, lg in zip(procs, logs):
            assert p.returncode == 0, lg[-3000:]
        return [dict(np.load(out + f".rank{r}.npz")) for r in range(world)]


def single_gpu(N, P, kernel):
    import gpax_b200
    X, y, Xn, theta = problem(N, P, kernel)
    return gpax_b200.default_context().posterior(kernel, X, y, Xn, theta[None], want=("mean", "var"))


@pytest.mark.parametrize("N,P,nb,kernel", [(2048, 300, 256, "Matern"), (3072, 100, 512, "RBF"), (1024, 700, 128, "Periodic")])
def test_one_rank_grid_matches_single_gpu(N, P, nb, kernel):
    ref = single_gpu(N, P, kernel)
    res = run_grid(1, 1, N, P, nb, kernel)[0]
    assert res["info"] == 0
    assert_close(res["mean"], ref["mean"][0], 1e-11, "mean, 1 x 1 grid")
    assert_close(res["var"], ref["var"][0], 1e-11, "var, 1 x 1 grid")


@pytest.mark.parametrize("pr,pc", [(1, 2), (2, 1), (2, 2), (2, 4)])
def test_process_grid_matches_single_gpu(pr, pc):
    if n_gpus() < pr * pc:
        pytest.skip(f"needs {pr * pc} GPUs")
    N, P, nb, kernel = 4096, 600, 256, "Matern"
    ref = single_gpu(N, P, kernel)
    res = run_grid(pr, pc, N, P, nb, kernel)
    for r in res:
        assert r["info"] == 0
        assert_close(r["mean"], ref["mean"][0], 1e-11, f"mean, {pr} x {pc} grid")      # equal to the 1-GPU result to 1e-12 .. 1e-11
        assert_close(r["var"], ref["var"][0], 1e-11, f"var, {pr} x {pc} grid")
        np.testing.assert_array_equal(r["mean"], res[0]["mean"])                       # every rank holds the same result
