"""Pins the oracle's linear-assignment restatement (oracle/gsm_oracle_impl.h ORC(lsa);
scipy's rectangular_lsap, which the reference calls per step for the polygon/line tasks —
readme.md:89-90, scipy==1.7.3 in requirements.txt:101) to scipy: committed golden vectors
from this image's scipy plus a live comparison."""
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle import gsm_oracle as O
from tests._util import GOLDEN


def test_golden_vectors_f64():
    z = np.load(os.path.join(GOLDEN, "lsa_scipy.npz"))
    n = int(z["count"])
    assert n >= 600
    for k in range(n):
        c = z[f"cost_{k}"]
        got = O.lsa(c)
        assert (got == z[f"col4row_{k}"]).all(), (k, c.shape)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 8, 12, 16, 32])
def test_live_scipy_random_and_ties(n):
    rng = np.random.default_rng(n)
    for t in range(120):
        c = rng.random((n, n)) if t % 3 == 0 else rng.integers(0, 2 + t % 4, (n, n)).astype(float)
        _, cols = linear_sum_assignment(c)
        assert (O.lsa(c) == cols).all()


def test_f32_matches_scipy_on_f32_costs():
    rng = np.random.default_rng(5)
    for n in (3, 6, 12):
        for _ in range(50):
            c = rng.random((n, n)).astype(np.float32)
            got = O.lsa(c)
            _, cols = linear_sum_assignment(c.astype(np.float64))
            # optimal value must agree; the permutation too unless fp32 rounding made a tie
            assert np.isclose(c[np.arange(n), got].sum(), c[np.arange(n), cols].sum(), rtol=1e-5)


def test_is_permutation_and_optimal_small():
    import itertools
    rng = np.random.default_rng(9)
    for n in (2, 3, 4, 5):
        for _ in range(30):
            c = rng.random((n, n))
            got = O.lsa(c)
            assert sorted(got) == list(range(n))
            best = min(sum(c[i, p[i]] for i in range(n)) for p in itertools.permutations(range(n)))
            assert np.isclose(c[np.arange(n), got].sum(), best)
