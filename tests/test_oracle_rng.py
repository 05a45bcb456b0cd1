"""Pin the oracle's numpy.RandomState recipes (oracle/mt19937.h) against numpy itself."""
import numpy as np
import pytest

import oracle

R = np.random.RandomState


@pytest.mark.parametrize('seed', [0, 1, 7, 327741615, 2 ** 32 - 1])
def test_rand_and_randint(seed):
  np.testing.assert_array_equal(oracle.rand(seed, 5000), R(seed).rand(5000))
  np.testing.assert_array_equal(oracle.randint(seed, 2 ** 32 - 1, 3000), R(seed).randint(2 ** 32 - 1, size=3000))
  np.testing.assert_array_equal(oracle.randint(seed, 3, 5000), R(seed).randint(0, 3, size=5000))
  np.testing.assert_array_equal(oracle.bits_i8(seed, 4099), R(seed).randint(2, size=4099, dtype='i1'))


def test_randint_i1_prefix_property():
  """The engine draws file-order bits for all candidates and uses the first K (K = kept count):
  a size-K call must be a prefix of a size-N call (illumina.py:93)."""
  a, b = R(5).randint(2, size=1001, dtype='i1'), R(5).randint(2, size=4000, dtype='i1')
  np.testing.assert_array_equal(a, b[:1001])


@pytest.mark.parametrize('p', [0.0125, 0.015, 0.025, 0.05, 0.09375, 0.1])
def test_geometric(p):
  n = 2000000
  np.testing.assert_array_equal(oracle.geometric(11, p, n), R(11).geometric(p, size=n))


@pytest.mark.parametrize('n', [1, 2, 3, 100, 65537, 300001])
def test_shuffle(n):
  x = np.arange(n, dtype=np.int64) * 3 + 1
  y = x.copy(); R(9).shuffle(y)
  np.testing.assert_array_equal(oracle.shuffle_i64(9, x), y)
  # shuffling a Python list (get_data_for_workers, readgenerate.py:156) consumes the same draws
  if n <= 100:
    z = list(x); R(9).shuffle(z)
    assert z == y.tolist()
