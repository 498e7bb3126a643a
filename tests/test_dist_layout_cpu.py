"""CPU: host-side logic of the in-library multi-GPU path -- the block-cyclic index algebra exported by libb200gp.so
(b2gp_dist_layout, no GPU needed) and the two-process TCP hand-over of the NCCL id (gpax_b200/dist.py)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from gpax_b200 import dist


@pytest.mark.parametrize("pr,pc", [(1, 1), (2, 4), (4, 2), (2, 2), (1, 2), (3, 2)])
def test_block_cyclic_tiles_partition_the_matrix(pr, pc):
    T, R, nb = 13, 3, 256
    owned = np.zeros((T + R, T), dtype=int)
    for r in range(pr):
        for c in range(pc):
            lr, lc, *_ = dist.layout(T, R, nb, pr, pc, r, c, 0)
            rows = [gi for gi in range(T + R) if gi % pr == r]
            cols = [gj for gj in range(T) if gj % pc == c]
            assert lr == len(rows) and lc == len(cols)
            for gi in rows:
                for gj in cols:
                    owned[gi, gj] += 1
    assert (owned == 1).all()


@pytest.mark.parametrize("pr,pc", [(2, 4), (4, 2), (3, 2)])
def test_panel_slots_hold_every_row_below_the_diagonal(pr, pc):
    T, R, nb = 11, 2, 128
    for k in range(T):
        total, slots = 0, set()
        for r in range(pr):
            _, _, prow, slot, first, _ = dist.layout(T, R, nb, pr, pc, r, 0, k)
            below = [gi for gi in range(T + R) if gi % pr == r and gi > k]
            assert prow == len(below) * nb
            assert first == len([gi for gi in range(T + R) if gi % pr == r and gi <= k])
            assert prow <= slot
            total += prow
            slots.add(slot)
        assert total == (T + R - k - 1) * nb and len(slots) == 1
        for c in range(pc):
            *_, fc = dist.layout(T, R, nb, pr, pc, 0, c, k)
            assert fc == len([gj for gj in range(T) if gj % pc == c and gj <= k])


def _id_worker(rank, world, port, q):
    class FakeLib:
        def b2gp_dist_unique_id(self, buf):
            buf.raw = bytes(range(128))
            return 0
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    q.put((rank, dist.exchange_id(FakeLib(), rank, world, port=port)))


def test_id_exchange_two_processes():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_id_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=60) for _ in range(3))
    for p in procs:
        p.join(timeout=30)
    assert all(got[r] == bytes(range(128)) for r in range(3))


def test_default_grid():
    assert dist.default_grid(8) == (2, 4) and dist.default_grid(4) == (2, 2) and dist.default_grid(2) == (1, 2)
    assert dist.default_grid(1) == (1, 1) and dist.default_grid(6) == (2, 3)
