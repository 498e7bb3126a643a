"""
distributed.py -- ROUND-1 PROTOTYPE of the two multi-GPU forms of the path, kept as the host-side reference of the 1-D
(block-column) layout and for its gloo CPU tests.  The product is the in-library version: gpax_b200/dist.py over
gpax_b200/csrc/dist.cuh (2-D block-cyclic layout, NCCL row / column communicators inside libb200gp.so, no torch).

The two forms that need a real exchange step:

  * `BlockCyclicGP`: exact-GP posterior with the N x N matrix distributed block-column-cyclically over
    the ranks (one process per GPU).  Right-looking Cholesky: the owner of block column k factors its
    diagonal block and solves its panel (C-ABI b2gp_potrf_inv / b2gp_trsm_inv), the panel is BROADCAST
    over NCCL, every rank applies the trailing update to the block columns it owns (b2gp_gemm_nt on the
    DMMA tensor pipe).  K is generated in place by the Gram kernel on the owning GPU -- it never exists
    anywhere else.  The solve is a fan-in forward substitution: every rank folds its solved blocks into
    the next block (GEMMs in parallel on all ranks), one small REDUCE per block, the owner finishes it.
    Replaces gpax/models/gp.py:253-277 for N beyond one GPU (BASELINE.json config 4).
  * `sharded_sparse_posterior`: viSparseGP posterior with the training set sharded over ranks: per-shard
    M x M and M statistics (b2gp_sparse_partial), one ALL-REDUCE, replicated finish (b2gp_sparse_finish).
    Replaces gpax/models/sparse_gp.py:173-223 (BASELINE.json config 5).

The draw-parallel mode (independent posteriors per rank, no exchange) needs none of this: see bench.py.

torch is used here for what the task brief allows it for: device buffers that NCCL can address and
`torch.distributed` collectives.  Every flop is in libb200gp.so.  `ops` is the local-compute interface:
`GpuOps` below (the product), or a NumPy stand-in injected by the CPU gloo tests (tests/test_distributed_cpu.py).
"""
import math

import numpy as np

from . import _ffi


def _cdiv(a, b):
    return (a + b - 1) // b


class GpuOps:
    """Local compute on torch CUDA fp64 tensors through the C-ABI (device-pointer mode)."""

    def __init__(self, ctx=None, device=None):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.ctx = ctx or _ffi.Context(self.device.index)
        self.lib, self.h = self.ctx.lib, self.ctx.h

    # ---- buffers
    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float64, device=self.device)

    # torch produces these buffers on ITS current stream; libb200gp consumes them on its own non-blocking streams, and
    # nothing else orders the two: wait for torch's stream before handing a freshly written buffer to the library
    def _ordered(self, t):
        self.torch.cuda.current_stream(self.device).synchronize()
        return t

    def zeros(self, shape):
        return self._ordered(self.torch.zeros(shape, dtype=self.torch.float64, device=self.device))

    def from_numpy(self, a):
        return self._ordered(self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device))

    def to_numpy(self, t):
        return t.detach().cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize(self.device)

    @staticmethod
    def _p(t):
        assert t.stride(-1) == 1 or t.numel() <= 1, "innermost dimension must be contiguous"
        return t.data_ptr()

    @staticmethod
    def _ld(t):
        return t.stride(0) if t.dim() == 2 and t.shape[0] > 1 else max(t.shape[-1], 1)

    # ---- compute
    def gram(self, kind, X, Z, theta, diag_add, same, out):
        d = X.shape[1]
        ell = np.ascontiguousarray(theta[:d], dtype=np.float64)
        self.ctx._check(self.lib.b2gp_gram(self.h, _ffi.KIND[kind], self._p(X), X.shape[0], self._p(Z), Z.shape[0], d,
                                           ell.ctypes.data, float(theta[d]), float(theta[d + 2]), float(diag_add),
                                           int(bool(same)), self._p(out), self._ld(out), _ffi.FLAG_DEVICE_PTRS))

    def potrf_inv(self, A, linv):
        import ctypes as C
        info = C.c_int(0)
        self.ctx._check(self.lib.b2gp_potrf_inv(self.h, A.shape[0], self._p(A), self._ld(A), self._p(linv), C.byref(info)))
        return info.value

    def trsm_inv(self, L, linv, B):
        if B.shape[0] == 0:
            return
        self.ctx._check(self.lib.b2gp_trsm_inv(self.h, L.shape[0], B.shape[0], self._p(L), self._ld(L), self._p(linv),
                                               self._p(B), self._ld(B)))

    def gemm_nt(self, A, B, Cm, alpha, beta, lower=False):
        m, k = A.shape
        n = B.shape[0]
        if m == 0 or n == 0:
            return
        self.ctx._check(self.lib.b2gp_gemm_nt(self.h, m, n, k, float(alpha), self._p(A), self._ld(A), self._p(B), self._ld(B),
                                              float(beta), self._p(Cm), self._ld(Cm), int(lower), _ffi.FLAG_DEVICE_PTRS))

    def rowdot(self, R, w, dot, nrm, accumulate):
        self.ctx._check(self.lib.b2gp_rowdot(self.h, R.shape[0], R.shape[1], self._p(R), self._ld(R),
                                             None if w is None else self._p(w), 1.0,
                                             None if dot is None else self._p(dot), None if nrm is None else self._p(nrm),
                                             int(accumulate)))

    def copy(self, dst, src):
        rows, cols = dst.shape
        self.ctx._check(self.lib.b2gp_copy2d(self.h, self._p(dst), self._ld(dst), self._p(src), self._ld(src), rows, cols))

    def sparse_partial(self, kind, Xu, Xtr, y, theta, jitter, Kpart, cpart):
        import ctypes as C
        info = C.c_int(0)
        th = np.ascontiguousarray(theta, dtype=np.float64)
        self.ctx._check(self.lib.b2gp_sparse_partial(self.h, _ffi.KIND[kind], self._p(Xu), Xu.shape[0], self._p(Xtr),
                                                     Xtr.shape[0], self._p(y), Xu.shape[1], th.ctypes.data, float(jitter),
                                                     self._p(Kpart), self._ld(Kpart), self._p(cpart), C.byref(info)))
        return info.value

    def sparse_finish(self, kind, Xu, Ksum, csum, Xnew, theta, noiseless, jitter, mean, var, cov):
        import ctypes as C
        info = C.c_int(0)
        th = np.ascontiguousarray(theta, dtype=np.float64)
        flags = _ffi.OUT_MEAN | (_ffi.OUT_VAR if var is not None else 0) | (_ffi.OUT_COV if cov is not None else 0)
        self.ctx._check(self.lib.b2gp_sparse_finish(self.h, _ffi.KIND[kind], self._p(Xu), Xu.shape[0], self._p(Ksum),
                                                    self._ld(Ksum), self._p(csum), self._p(Xnew), Xnew.shape[0], Xu.shape[1],
                                                    th.ctypes.data, int(bool(noiseless)), float(jitter), flags, self._p(mean),
                                                    None if var is None else self._p(var),
                                                    None if cov is None else self._p(cov), C.byref(info)))
        return info.value


def kdiag_value(kind, theta, d):
    """k(x, x) exactly as the Gram kernel evaluates a zero distance (gram.cuh: cov_self)."""
    scale = float(theta[d])
    if kind == "Matern":
        r = math.sqrt(0.0 + 1e-12)
        s5r = 5 ** 0.5 * r
        return scale * (1 + s5r + (5 / 3) * 0.0) * math.exp(-s5r)
    return scale


class BlockCyclicGP:
    """Exact-GP posterior (mean + diagonal variance) with K distributed by block columns, cyclically."""

    def __init__(self, ops, N, nb, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.ops, self.N, self.nb = ops, int(N), int(nb)
        assert self.nb % 128 == 0, "block width must be a multiple of the 128-wide leaf"
        self.nblk = _cdiv(self.N, self.nb)
        self.owned = [j for j in range(self.nblk) if j % self.world == self.rank]
        self.col = {j: ops.empty((self.N - j * self.nb, self.nb)) for j in self.owned}
        self.linv = {j: ops.empty((_cdiv(self.width(j), 128) * 128 * 128,)) for j in self.owned}
        self.pbuf = ops.empty((self.N, self.nb))
        self.info = 0
        self.bytes_broadcast = 0

    def width(self, j):
        return min(self.nb, self.N - j * self.nb)

    def owner(self, j):
        return j % self.world

    def _sync(self):
        self.ops.sync()

    # ---- K generated in place on the owning rank (gp.py:269: kernel(X_train, X_train, params, noise, jitter))
    def build(self, kind, X, theta, jitter):
        d = X.shape[1]
        diag = float(theta[d + 1]) + float(jitter)
        for j in self.owned:
            w = self.width(j)
            self.ops.gram(kind, X[j * self.nb:], X[j * self.nb: j * self.nb + w], theta, diag, True, self.col[j][:, :w])

    # ---- right-looking factorisation with look-ahead: while the ranks apply panel k to the block columns they
    #      own, the owner of column k+1 has already brought that column up to date, factored it, and its panel
    #      is travelling (asynchronous NCCL broadcast into the second panel buffer)
    def _panel(self, k, info):
        """owner only: factor the diagonal block of column k and solve the rows below it"""
        w, rows = self.width(k), self.N - k * self.nb
        panel = self.col[k]
        i = self.ops.potrf_inv(panel[:w, :w], self.linv[k])
        if i and not info[0]:
            info[0] = k * self.nb + i
        if rows > w:
            self.ops.trsm_inv(panel[:w, :w], self.linv[k], panel[w:, :w])

    def _update(self, panel, k, j):
        """col[j] -= panel[rows of j] * panel[block j]^T with panel = L[k*nb:, block k]"""
        off, wj, w = (j - k) * self.nb, self.width(j), self.width(k)
        self.ops.gemm_nt(panel[off:, :w], panel[off: off + wj, :w], self.col[j][:, :wj], -1.0, 1.0)

    def _bcast(self, k, buf, async_op):
        """broadcast panel k from its owner; returns (view of the panel on this rank, work handle or None)"""
        rows, own = self.N - k * self.nb, self.owner(k)
        panel = self.col[k] if self.rank == own else buf[:rows]
        work = None
        if self.world > 1:
            self._sync()
            src = self.dist.get_global_rank(self.group, own) if self.group else own
            work = self.dist.broadcast(panel, src=src, group=self.group, async_op=async_op)
            self.bytes_broadcast += panel.numel() * 8
            if not async_op:
                self._sync()
                work = None
        return panel, work

    def factor(self):
        info = [0]
        if not hasattr(self, "pbuf2"):
            self.pbuf2 = self.ops.empty((self.N, self.nb)) if self.world > 1 else self.pbuf
        cur, nxt = self.pbuf, self.pbuf2
        if self.rank == self.owner(0):
            self._panel(0, info)
        panel, _ = self._bcast(0, cur, False)
        for k in range(self.nblk):
            nxt_panel, work = None, None
            if k + 1 < self.nblk:
                if self.rank == self.owner(k + 1):
                    self._update(panel, k, k + 1)          # look-ahead column first ...
                    self._panel(k + 1, info)               # ... factor it ...
                nxt_panel, work = self._bcast(k + 1, nxt, True)   # ... and send it while the rest is updated
            for j in self.owned:
                if j > k + 1:
                    self._update(panel, k, j)
            if work is not None:
                work.wait()
                self._sync()
            panel = nxt_panel
            cur, nxt = nxt, cur
        self.info = self._max_int(info[0])
        return self.info

    def _max_int(self, v):
        """first failing pivot over all ranks (0 if none)"""
        if self.world == 1:
            return v
        import torch
        t = torch.tensor([1 if v else 0, -v if v else -(2 ** 40)], dtype=torch.int64, device=self.pbuf.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(-t[1].item()) if int(t[0].item()) else 0

    # ---- fan-in forward substitution + mean / diagonal-variance epilogue (gp.py:268, 272-273; vigp.py:185)
    def solve_mean_var(self, kind, X, y, Xnew, theta, noiseless, jitter):
        ops, nb, N = self.ops, self.nb, self.N
        P, d = Xnew.shape[0], X.shape[1]
        P1 = P + 1
        stage = ops.empty((P1, nb))
        zero = ops.zeros((P1, nb))
        xloc = {}
        for k in range(self.nblk):
            w, own = self.width(k), self.owner(k)
            st = stage[:, :w]
            first = True
            if self.rank == own:
                # [k_pX block ; y block]: kernel(X_new, X_train, params, jitter=0.0) restricted to this block
                ops.gram(kind, Xnew, X[k * nb: k * nb + w], theta, 0.0, False, stage[:P, :w])
                ops.copy(stage[P:P1, :w], y[k * nb: k * nb + w].reshape(1, w))
                first = False
            for j in self.owned:
                if j >= k:
                    continue
                off = (k - j) * nb
                ops.gemm_nt(xloc[j], self.col[j][off: off + w, :self.width(j)], st, -1.0, 0.0 if first else 1.0)
                first = False
            if first:
                ops.copy(st, zero[:, :w])
            if self.world > 1:
                self._sync()
                contig = stage if w == nb else st.contiguous()
                self.dist.reduce(contig, dst=self.dist.get_global_rank(self.group, own) if self.group else own,
                                 op=self.dist.ReduceOp.SUM, group=self.group)
                self._sync()
                if w != nb and self.rank == own:
                    ops.copy(st, contig)
            if self.rank == own:
                ops.trsm_inv(self.col[k][:w, :w], self.linv[k], st)
                xk = ops.empty((P1, w))
                ops.copy(xk, st)
                xloc[k] = xk
        mean = ops.zeros((P,))
        sq = ops.zeros((P,))
        for k in self.owned:
            xk = xloc[k]
            ops.rowdot(xk[:P], xk[P], mean, sq, True)
        if self.world > 1:
            self._sync()
            self.dist.all_reduce(mean, op=self.dist.ReduceOp.SUM, group=self.group)
            self.dist.all_reduce(sq, op=self.dist.ReduceOp.SUM, group=self.group)
            self._sync()
        mean_h, sq_h = ops.to_numpy(mean), ops.to_numpy(sq)
        kd = kdiag_value(kind, theta, d) + (float(theta[d + 1]) * (0.0 if noiseless else 1.0) + float(jitter))
        var_h = kd - sq_h
        if self.info:
            mean_h = np.full_like(mean_h, np.nan)
            var_h = np.full_like(var_h, np.nan)
        return mean_h, var_h

    def posterior(self, kind, X, y, Xnew, theta, noiseless=False, jitter=1e-6):
        """X [N,d], y [N], Xnew [P,d]: device tensors replicated on every rank (a few hundred KB);
        theta: host array (d+3).  Returns (mean[P], var[P], info) as NumPy on every rank."""
        self.build(kind, X, theta, jitter)
        self.factor()
        mean, var = self.solve_mean_var(kind, X, y, Xnew, theta, noiseless, jitter)
        return mean, var, self.info


def sharded_sparse_posterior(ops, kind, Xu, Xtr_shard, y_shard, Xnew, theta, noiseless=False, jitter=1e-6,
                             want_cov=False, group=None):
    """viSparseGP posterior with the training set sharded over the ranks of `group`.
    Xu [M,d], Xnew [P,d] replicated; Xtr_shard [N_r,d], y_shard [N_r] local.  One all-reduce of M*M + M doubles."""
    import torch.distributed as dist
    M, P = Xu.shape[0], Xnew.shape[0]
    Kpart, cpart = ops.zeros((M, M)), ops.zeros((M,))
    info = ops.sparse_partial(kind, Xu, Xtr_shard, y_shard, theta, jitter, Kpart, cpart)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        ops.sync()
        dist.all_reduce(Kpart, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(cpart, op=dist.ReduceOp.SUM, group=group)
        ops.sync()
    mean, var = ops.empty((P,)), ops.empty((P,))
    cov = ops.empty((P, P)) if want_cov else None
    info2 = ops.sparse_finish(kind, Xu, Kpart, cpart, Xnew, theta, noiseless, jitter, mean, var, cov)
    out = {"mean": ops.to_numpy(mean), "var": ops.to_numpy(var), "info": info or info2}
    if want_cov:
        out["cov"] = ops.to_numpy(cov)
    return out
