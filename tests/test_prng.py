"""The host PRNG (gpax_b200/prng.py) against the known answers JAX publishes for its threefry stream.

JAX cannot be installed here, so these constants are the pin:
  * the Threefry-2x32 known-answer vectors of the Random123 distribution, which JAX's own test-suite checks
    (tests/random_test.py::testThreefry2x32);
  * `jax.random.split(jax.random.PRNGKey(0))`, `jax.random.uniform(PRNGKey(0), (1,))`, `jax.random.normal(PRNGKey(0), (1,))`
    and `jax.random.normal(PRNGKey(42), (3,))` as printed in the JAX documentation (JAX 0.4.x, the generation the
    reference was written against: original, "non-partitionable" threefry layout).
"""
import numpy as np
import pytest

from gpax_b200 import prng, utils


@pytest.mark.parametrize("key,count,expected", [
    ((0x0, 0x0), (0x0, 0x0), (0x6b200159, 0x99ba4efe)),
    ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
    ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0)),
])
def test_threefry2x32_known_answers(key, count, expected):
    out = prng._threefry_2x32(np.array(key, np.uint32), np.array(count, np.uint32))
    assert tuple(int(v) for v in out) == expected


def test_key_and_split_match_jax_docs():
    np.testing.assert_array_equal(prng.PRNGKey(0), np.array([0, 0], np.uint32))
    np.testing.assert_array_equal(prng.PRNGKey(2**32 + 5), np.array([1, 5], np.uint32))
    np.testing.assert_array_equal(prng.split(prng.PRNGKey(0)), np.array([[4146024105, 967050713], [2718843009, 1272950319]], np.uint32))
    k1, k2 = utils.get_keys(0)                                  # gpax/utils/utils.py:24-30
    np.testing.assert_array_equal(k1, [4146024105, 967050713])
    np.testing.assert_array_equal(k2, [2718843009, 1272950319])


def test_uniform_and_normal_match_jax_docs():
    np.testing.assert_allclose(prng.uniform(prng.PRNGKey(0), (1,)), [0.41845703], rtol=0, atol=1e-7)
    np.testing.assert_allclose(prng.normal(prng.PRNGKey(0), (1,)), [-0.20584226], rtol=0, atol=1e-7)
    np.testing.assert_allclose(prng.normal(prng.PRNGKey(42), (3,)), [0.18693547, -1.2806505, -1.5593132], rtol=0, atol=5e-7)


def test_odd_sizes_64_bit_words_and_partitionable_layout():
    key = prng.PRNGKey(7)
    a = prng.random_bits(key, 32, (5,))
    b = prng.random_bits(key, 32, (6,))
    assert a.dtype == np.uint32 and a.shape == (5,) and not np.array_equal(a, b[:5])   # halves move with the size
    w = prng.random_bits(key, 64, (3, 2))
    assert w.dtype == np.uint64 and w.shape == (3, 2)
    raw = prng._threefry_2x32(key, np.arange(12, dtype=np.uint32))
    np.testing.assert_array_equal(w.ravel(), (raw[:6].astype(np.uint64) << np.uint64(32)) | raw[6:].astype(np.uint64))
    x = prng.normal(key, (4000,), np.float64)
    assert x.dtype == np.float64 and abs(x.mean()) < 0.06 and abs(x.std() - 1.0) < 0.05
    p = prng.random_bits(key, 32, (2, 3), partitionable=True)
    hi, lo = prng.threefry2x32(key[0], key[1], np.zeros(6, np.uint32), np.arange(6, dtype=np.uint32))
    np.testing.assert_array_equal(p.ravel(), hi ^ lo)
    assert prng.split(key, 3, partitionable=True).shape == (3, 2)


def test_posterior_eps_follows_the_reference_key_discipline():
    key = prng.PRNGKey(3)
    eps = utils.posterior_eps(key, 4, 2, 5)                     # predict: split per draw (gp.py:391), float32 variates
    keys = prng.split(key, 4)
    for s in range(4):
        np.testing.assert_array_equal(eps[s], prng.normal(keys[s], (2, 5), np.float32).astype(np.float64))
    one = utils.posterior_eps(key, 1, 2, 5, per_draw_keys=False)    # _predict: the key as it is (gp.py:292)
    np.testing.assert_array_equal(one[0], prng.normal(key, (2, 5), np.float32).astype(np.float64))
    np.testing.assert_array_equal(utils.posterior_eps(3, 4, 2, 5), eps)          # an int seed is PRNGKey(seed)
    e64 = utils.posterior_eps(key, 4, 2, 5, np.float64)
    assert not np.array_equal(e64, eps)                             # x64 draws come from 64-bit words
    g = utils.posterior_eps(np.random.default_rng(1), 2, 1, 3)
    np.testing.assert_array_equal(g, np.random.default_rng(1).standard_normal((2, 1, 3)))
    with pytest.raises(TypeError):
        prng.as_key(np.array([1.0, 2.0]))
