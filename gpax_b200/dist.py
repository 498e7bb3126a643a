"""
dist.py -- host side of the in-library multi-GPU path (gpax_b200/csrc/dist.cuh): one process per GPU, rank / world /
rendezvous from the launcher's environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT as torchrun sets
them).  Python only carries the 128-byte NCCL id from rank 0 to the other ranks over a TCP socket and calls the
collective C-ABI entry points; the exchange steps themselves (panel broadcasts, all-gathers, the M x M all-reduce) are
NCCL calls issued inside libb200gp.so on its own streams.  No torch.
"""
import ctypes as C
import importlib.util
import os
import socket
import time

import numpy as np

from . import _ffi


def _find_nccl():
    """libnccl.so.2: B200GP_NCCL_LIB if set, else the wheel torch ships (nvidia/nccl/lib), else the system loader's."""
    if os.environ.get("B200GP_NCCL_LIB"):
        return
    try:
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            p = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(p):
                os.environ["B200GP_NCCL_LIB"] = p
                return
    except Exception:  # noqa: BLE001
        pass


def exchange_id(lib, rank, world, addr=None, port=None, timeout=120.0):
    """rank 0 creates the NCCL unique id and serves it to the other world-1 ranks; returns the 128 bytes on every rank"""
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(port or os.environ.get("B200GP_DIST_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 17))
    if rank == 0:
        buf = C.create_string_buffer(128)
        rc = lib.b2gp_dist_unique_id(buf)
        if rc != 0:
            raise _ffi.B200GPError(f"b2gp_dist_unique_id failed ({rc}): libnccl.so.2 not loadable; set B200GP_NCCL_LIB")
        ident = buf.raw
        if world > 1:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(world)
            srv.settimeout(timeout)
            for _ in range(world - 1):
                conn, _ = srv.accept()
                conn.sendall(ident)
                conn.close()
            srv.close()
        return ident
    deadline = time.time() + timeout
    while True:
        try:
            s = socket.create_connection((addr, port), timeout=5.0)
            break
        except OSError:
            if time.time() > deadline:
                raise
            time.sleep(0.2)
    ident = b""
    while len(ident) < 128:
        chunk = s.recv(128 - len(ident))
        if not chunk:
            raise _ffi.B200GPError("rank 0 closed the id socket early")
        ident += chunk
    s.close()
    return ident


def default_grid(world):
    """2 x 4 for 8 GPUs, otherwise the most square pr <= pc factorisation"""
    pr = int(np.floor(np.sqrt(world)))
    while world % pr:
        pr -= 1
    return pr, world // pr


class DistContext:
    """A libb200gp context that is one member of a process grid.  `ctx` defaults to a new Context on LOCAL_RANK."""

    def __init__(self, ctx=None, grid=None, rank=None, world=None):
        _find_nccl()
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
        self.ctx = ctx or _ffi.Context(int(os.environ.get("LOCAL_RANK", "0")))
        self.grid = tuple(grid) if grid else default_grid(self.world)
        if self.grid[0] * self.grid[1] != self.world:
            raise ValueError(f"grid {self.grid} does not cover {self.world} ranks")
        ident = exchange_id(self.ctx.lib, self.rank, self.world)
        self.ctx._check(self.ctx.lib.b2gp_dist_init(self.ctx.h, ident, self.rank, self.world, self.grid[0], self.grid[1]))

    def close(self):
        if self.ctx is not None and self.ctx.h is not None:
            self.ctx.lib.b2gp_dist_finalize(self.ctx.h)

    def posterior(self, kind, Xtr, yres, Xnew, theta, noiseless=False, jitter=1e-6, nb=512, want_var=True):
        """COLLECTIVE.  Exact-GP posterior mean / diagonal variance with k_XX block-cyclic over the grid (one theta)."""
        Xtr, Xnew = _ffi._f64(Xtr), _ffi._f64(Xnew)
        N, d = Xtr.shape
        P = Xnew.shape[0]
        theta = _ffi._f64(theta).reshape(d + 3)
        yres = _ffi._f64(yres).reshape(N)
        mean, var = np.empty(P), (np.empty(P) if want_var else None)
        info = C.c_int(0)
        t = _ffi.Timing()
        flags = _ffi.OUT_MEAN | (_ffi.OUT_VAR if want_var else 0)
        self.ctx._check(self.ctx.lib.b2gp_dist_posterior(
            self.ctx.h, _ffi.KIND[kind] if isinstance(kind, str) else kind, _ffi._ptr(Xtr), N, _ffi._ptr(yres), _ffi._ptr(Xnew), P, d,
            _ffi._ptr(theta), int(bool(noiseless)), float(jitter), int(nb), flags, _ffi._ptr(mean), _ffi._ptr(var), C.byref(info),
            C.byref(t)))
        return {"mean": mean, "var": var, "info": info.value, "timing": t.as_dict()}


def _sparse_posterior(self, kind, Xu, Xtr_shard, y_shard, Xnew, theta, noiseless=False, jitter=1e-6, want_var=True):
    """COLLECTIVE.  N-sharded sparse posterior: this rank's shard of (X, y), the same Xu / X_new / theta on every rank."""
    Xu, Xs, Xnew = _ffi._f64(Xu), _ffi._f64(Xtr_shard), _ffi._f64(Xnew)
    M, d = Xu.shape
    Ns, P = Xs.shape[0], Xnew.shape[0]
    theta = _ffi._f64(theta).reshape(d + 3)
    ys = _ffi._f64(y_shard).reshape(Ns)
    mean, var = np.empty(P), (np.empty(P) if want_var else None)
    info = C.c_int(0)
    t = _ffi.Timing()
    flags = _ffi.OUT_MEAN | (_ffi.OUT_VAR if want_var else 0)
    self.ctx._check(self.ctx.lib.b2gp_dist_sparse_posterior(
        self.ctx.h, _ffi.KIND[kind] if isinstance(kind, str) else kind, _ffi._ptr(Xu), M, _ffi._ptr(Xs), Ns, _ffi._ptr(ys), _ffi._ptr(Xnew),
        P, d, _ffi._ptr(theta), int(bool(noiseless)), float(jitter), flags, _ffi._ptr(mean), _ffi._ptr(var), C.byref(info), C.byref(t)))
    return {"mean": mean, "var": var, "info": info.value, "timing": t.as_dict()}


DistContext.sparse_posterior = _sparse_posterior


def layout(T, R, nb, pr, pc, row, col, k):
    """b2gp_dist_layout: (local tile rows, local tile cols, panel rows at step k, slot rows, first row after k, first col after k)"""
    lib = _ffi.load_library()
    out = (C.c_int64 * 6)()
    rc = lib.b2gp_dist_layout(T, R, nb, pr, pc, row, col, k, out)
    if rc != 0:
        raise ValueError("bad layout arguments")
    return tuple(out)
