"""threefry2x32 / jax.random restatement (flaxdiff_b200/utils.py) against PUBLISHED vectors:
Random123 known-answer tests for Threefry-2x32-20 and the JAX documentation examples for
PRNGKey(0).  This pins FourierEmbedding.freqs = normal(PRNGKey(42), (128,)) * 16
(flaxdiff/models/common.py:101-102), the only RNG-derived constant of the model."""
import numpy as np

from flaxdiff_b200 import utils as u


def _tf(key, ctr):
    a, b = u.threefry2x32(key, np.array([ctr[0]], dtype=np.uint32), np.array([ctr[1]], dtype=np.uint32))
    return int(a[0]), int(b[0])


def test_threefry_random123_kat():
    assert _tf((0, 0), (0, 0)) == (0x6B200159, 0x99BA4EFE)
    assert _tf((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF)) == (0x1CB996FC, 0xBB002BE7)
    assert _tf((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3)) == (0xC4923A9C, 0x483DF7A0)


def test_split_matches_jax_docs():
    # jax.random.split(jax.random.PRNGKey(0)) in the JAX PRNG design notes
    assert u.split(u.PRNGKey(0)) == [(4146024105, 967050713), (2718843009, 1272950319)]


def test_normal_matches_jax_quickstart():
    # jax.random.normal(jax.random.PRNGKey(0), (10,)) printed in the JAX quick-start
    want = np.array([-0.3721109, 0.26423115, -0.18252768, -0.7368197, -0.44030377, -0.1521442,
                     -0.67135346, -0.5908641, 0.73168886, 0.5673026], dtype=np.float32)
    got = u.normal(u.PRNGKey(0), 10)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-7)


def test_fourier_freqs_constant():
    f = u.normal(u.PRNGKey(42), 128) * 16.0
    assert f.shape == (128,) and f.dtype == np.float32
    assert abs(float(f.std()) - 16.0) < 3.0
    g = u.normal(u.PRNGKey(42), 128) * 16.0
    assert np.array_equal(f, g)


def test_random_markov_state_is_functional():
    s = u.RandomMarkovState(u.PRNGKey(7))
    s1, k1 = s.get_random_key()
    s2, k2 = s.get_random_key()
    assert k1 == k2 and s1 == s2 and k1 != s1.rng
    assert u.fold_in(k1, 0) != u.fold_in(k1, 1)
