"""The digit-plane (Ozaki) GEMM restated in NumPy (oracle/ozaki_oracle.py): the claims DESIGN.md 4.6 makes about the
int8 tcgen05 path, checked in exact arithmetic on the CPU, and the kernel itself checked against the restatement."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import ozaki_oracle as oz


def _rows(rng, m, k):
    return rng.standard_normal((m, k)) * np.exp(rng.normal(0, 2, (m, 1)))      # rows of very different scale


def test_digits_are_int8_and_reconstruct_to_2_pow_minus_55():
    rng = np.random.default_rng(0)
    A = _rows(rng, 7, 50)
    A[3] = 0.0                                                      # an all-zero row keeps exponent 0
    A[5, 0] = -A[5].__abs__().max() * 1.0                           # the row maximum itself, negative
    for S in (7, 8):
        planes, scale = oz.slice_rows(A, S)
        assert planes.min() >= -64 and planes.max() <= 64
        for i in range(A.shape[0]):
            for j in range(0, A.shape[1], 7):
                exact = Fraction(float(A[i, j]))
                got = sum(Fraction(int(planes[p, i, j]), 2 ** (6 + 7 * p)) for p in range(S)) * Fraction(float(scale[i]))
                assert abs(exact - got) <= Fraction(float(scale[i])) / 2 ** (6 + 7 * (S - 1) + 1)    # half a unit of the last digit
    assert scale[3] == 1.0


def test_class_sums_are_exact_integers_within_int32():
    rng = np.random.default_rng(1)
    A, B = _rows(rng, 5, 300), _rows(rng, 4, 300)
    PA, _ = oz.slice_rows(A, 8)
    PB, _ = oz.slice_rows(B, 8)
    D = oz.class_sums(PA, PB)
    for t in (0, 3, 7):
        ref = sum(int(PA[p, 2, kk]) * int(PB[t - p, 1, kk]) for p in range(t + 1) for kk in range(300))
        assert int(D[t, 2, 1]) == ref
    assert np.abs(D).max() < 2 ** 31
    assert 8 * 32768 * 64 * 64 == 2 ** 30                           # S pairs per class, k at the kernel's limit, digits at +-64:
                                                                    # the int32 bound the k <= 32768 check in ozaki.cuh relies on


@pytest.mark.parametrize("S,bound", [(8, 2.0 ** -50), (7, 2.0 ** -44)])
def test_gemm_error_against_exact_rational(S, bound):
    rng = np.random.default_rng(2)
    m, n, k = 6, 5, 96
    A, B, C = _rows(rng, m, k), _rows(rng, n, k), rng.standard_normal((m, n))
    out = oz.gemm_nt(A, B, C, alpha=-1.0, S=S)
    for i in range(m):
        for j in range(n):
            exact = Fraction(float(C[i, j])) - sum(Fraction(float(A[i, kk])) * Fraction(float(B[j, kk])) for kk in range(k))
            scale = float(np.linalg.norm(A[i]) * np.linalg.norm(B[j]) + abs(C[i, j]))
            assert abs(float(Fraction(float(out[i, j])) - exact)) <= bound * scale
    low = oz.gemm_nt(A[:5], A[:5], C[:5, :5], alpha=-1.0, S=S, lower_only=True)
    full = oz.gemm_nt(A[:5], A[:5], C[:5, :5], alpha=-1.0, S=S)
    np.testing.assert_array_equal(np.tril(low), np.tril(full))
    np.testing.assert_array_equal(np.triu(low, 1), np.triu(C[:5, :5], 1))


@pytest.mark.gpu
@pytest.mark.xfail(reason="written after the round's GPU budget was spent: first hardware run pending", strict=False)
@pytest.mark.parametrize("S", [8, 7])
def test_int8_kernel_equals_the_restatement_bit_for_bit(S):
    """exact integer products + the same fixed-order fp64 recombination on both sides -> identical doubles"""
    import gpax_b200
    ctx = gpax_b200.default_context()
    rng = np.random.default_rng(3)
    m, n, k = 1536, 1280, 544                                       # >= 148 tiles of 128 x 64, k >= 512, ragged k-block
    A, B, C = _rows(rng, m, k), _rows(rng, n, k), rng.standard_normal((m, n))
    try:
        ctx.set_option("ozaki", S)
        got = ctx.gemm_nt(A, B, C, alpha=-1.0, beta=1.0)
    finally:
        ctx.set_option("ozaki", 8)
    np.testing.assert_array_equal(got, oz.gemm_nt(A, B, C, alpha=-1.0, S=S))
