"""CPU: properties of the systematic-resampling oracle (oracle/resample.py) -- the definition the
integer kernel is bit-exact against."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle.resample import exp_det, fixed_point_weights, systematic_ancestors


def test_exp_det_matches_libm_to_1ulp_scale():
    t = -np.linspace(0, 59.9, 5000)
    assert np.max(np.abs(exp_det(t) / np.exp(t) - 1)) < 1e-14


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 300), st.integers(0, 2 ** 32 - 1), st.floats(0.1, 30.0), st.integers(0, 10 ** 6))
def test_ancestor_properties(n, u0, spread, seed):
    rng = np.random.default_rng(seed)
    lw = (rng.standard_normal(n) * spread).astype(np.float32)
    anc = systematic_ancestors(lw, u0)
    assert anc.shape == (n,) and anc.min() >= 0 and anc.max() < n
    assert np.all(np.diff(anc) >= 0)                                   # sorted
    q = fixed_point_weights(lw).astype(np.float64)
    counts = np.bincount(anc, minlength=n)
    assert np.all(np.abs(counts - n * q / q.sum()) < 1.0 + 1e-9)       # systematic: |N_i - n w_i| < 1
    assert np.all(counts[q == 0] == 0)                                 # zero-weight rows never chosen


def test_uniform_weights_give_identity_and_nonfinite_are_dropped():
    assert np.array_equal(systematic_ancestors(np.zeros(64, np.float32), 999), np.arange(64))
    lw = np.array([0.0, np.nan, 0.0, -np.inf, np.inf, 0.0], dtype=np.float32)
    anc = systematic_ancestors(lw, 2 ** 31)
    assert set(anc.tolist()) <= {0, 2, 5}
