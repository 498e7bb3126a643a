"""ozaki_oracle.py -- NumPy restatement of the int8 digit-plane GEMM of gpax_b200/csrc/ozaki.cuh.  TEST INFRASTRUCTURE ONLY.

The CUDA kernel evaluates  C += alpha * A B^T  (A [m,k], B [n,k], fp64) as S(S+1)/2 exact int8 x int8 -> int32 products
of digit planes plus a fixed-order fp64 recombination.  Every step of that is either exact integer arithmetic or a single
correctly rounded fp64 operation, so the whole computation can be restated on the CPU **bit for bit**:

  slice (oz_slice_kernel):   e_i = exponent of max_k |a_ik| (frexp), v = a * 2^(6 - e_i);
                             repeat S times: d = rint(v) (ties to even), v = (v - d) * 128        -- all exact in fp64
  products (oz_mma_kernel):  D_t = sum over digit pairs (p, q) with p + q = t of  A_p B_q^T        -- exact integers
  recombine (epilogue):      acc = 0; for t = S-1 .. 0: acc = fma(D_t, 2^-(12 + 7 t), acc)         -- D_t * 2^x is exact,
                                                                                                     one rounding per t
                             C = fma(2^e_i * alpha * 2^f_j, acc, C)                                -- exact product when
                                                                                                     alpha = +-1 (the only
                                                                                                     values the solver uses)
This module is the checker for that path (tests/test_ozaki_oracle.py pins its error bounds on the CPU; the GPU test
compares the kernel's output with it for equality).  Only tests/ may import it."""
import numpy as np


def slice_rows(A, S):
    """Digit planes [S, rows, k] (int64, values in [-64, 64]) and the row scales 2^e of oz_slice_kernel."""
    A = np.asarray(A, dtype=np.float64)
    mx = np.abs(A).max(axis=1) if A.shape[1] else np.zeros(A.shape[0])
    _, e = np.frexp(mx)                                   # mx = f * 2^e, f in [0.5, 1)
    e = np.where((mx > 0) & (mx < 1e300), e, 0).astype(np.int64)
    v = np.ldexp(A, (6 - e)[:, None].astype(np.int32))
    planes = np.empty((S,) + A.shape, dtype=np.int64)
    for p in range(S):
        d = np.rint(v)
        planes[p] = d.astype(np.int64)
        v = (v - d) * 128.0
    return planes, np.ldexp(1.0, e.astype(np.int32))


def class_sums(PA, PB):
    """D_t = sum_{p+q=t} A_p B_q^T for t < S, as exact int64 (the int32 TMEM accumulators hold the same values)."""
    S = PA.shape[0]
    D = np.zeros((S, PA.shape[1], PB.shape[1]), dtype=np.int64)
    FA, FB = PA.astype(np.float64), PB.astype(np.float64)     # |digit| <= 64, k <= 32768: every dot product is below 2^27,
    for p in range(S):                                        # so the float64 BLAS product is exact
        for q in range(S - p):
            D[p + q] += np.rint(FA[p] @ FB[q].T).astype(np.int64)
    return D


def gemm_nt(A, B, C, alpha=-1.0, S=8, lower_only=False):
    """C + alpha * A B^T exactly as the CUDA path rounds it.  alpha must be +-1 (power-of-two scale -> exact product)."""
    if abs(alpha) != 1.0:
        raise ValueError("bit-exact restatement needs alpha = +-1")
    A, B = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    if A.shape[1] > 32768:
        raise ValueError("k beyond the int32 accumulation bound of the kernel")
    PA, sa = slice_rows(A, S)
    PB, sb = slice_rows(B, S)
    D = class_sums(PA, PB)
    assert np.abs(D).max() < 2 ** 31, "int32 accumulator overflow"
    acc = np.zeros(D.shape[1:])
    for t in range(S - 1, -1, -1):
        acc = D[t].astype(np.float64) * np.ldexp(1.0, -(12 + 7 * t)) + acc      # exact product, one rounding: the fma
    out = (sa[:, None] * alpha * sb[None, :]) * acc + np.asarray(C, dtype=np.float64)
    if lower_only:
        out = np.where(np.tril(np.ones(out.shape, bool)), out, C)
    return out
