"""Oracle for systematic resampling with a fixed-point CDF.  TEST INFRASTRUCTURE.

The reference has no systematic resampler (only multinomial `resample`,
fab/sampling_methods/base.py:121-124); this is the build-side
extension BASELINE.json's north_star asks for, defined so that a parallel GPU scan is
bit-exact against this sequential restatement:

  m      = max_i log_w[i]                       over finite entries (fp32)
  t_i    = float64(fp32(log_w[i] - m))          non-finite log_w -> weight 0
  q_i    = floor(exp_det(t_i) * 2^30)           uint64; exp_det = IEEE-only double routine below
  c_i    = q_0 + ... + q_i                      uint64 inclusive scan (exact, associative)
  S      = c_{N-1}
  anc_k  = min{ i : c_i * (N * 2^32) > (k * 2^32 + u0) * S },  k = 0..N-1,  u0 in [0, 2^32)

i.e. positions (k + u0/2^32)/N against the normalised CDF, compared in exact 128-bit
integer arithmetic.  `exp_det` uses only IEEE-754 double add/mul/rint/ldexp (no fused
multiply-add, no libm exp), so numpy on the host and `__dmul_rn/__dadd_rn` on the device
give identical bits.
"""
import bisect

import numpy as np

LOG2E = float.fromhex("0x1.71547652b82fep+0")
LN2_HI = float.fromhex("0x1.62e42fee00000p-1")
LN2_LO = float.fromhex("0x1.a39ef35793c76p-33")
# 1/n! for n = 0..13
INV_FACT = [1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
            1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800]
WEIGHT_BITS = 30
T_MIN = -60.0   # exp(-60)*2^30 < 1  ->  weight 0


def exp_det(t: np.ndarray) -> np.ndarray:
    """Deterministic exp for t <= 0 (float64 in, float64 out)."""
    t = np.asarray(t, dtype=np.float64)
    t = np.maximum(t, T_MIN)
    k = np.rint(t * LOG2E)
    r = (t - k * LN2_HI) - k * LN2_LO
    p = np.full_like(r, INV_FACT[13])
    for n in range(12, -1, -1):
        p = p * r + INV_FACT[n]          # separate mul and add (numpy never fuses)
    return np.ldexp(p, k.astype(np.int64))


def fixed_point_weights(log_w: np.ndarray) -> np.ndarray:
    lw = np.asarray(log_w, dtype=np.float32)
    finite = np.isfinite(lw)
    if not finite.any():
        return np.zeros(lw.shape, dtype=np.uint64)
    m = np.float32(lw[finite].max())
    t = (np.where(finite, lw, m) - m).astype(np.float32).astype(np.float64)
    e = exp_det(t)
    q = np.floor(e * float(1 << WEIGHT_BITS)).astype(np.uint64)
    q[~finite] = 0
    q[t < T_MIN] = 0
    return q


def systematic_ancestors(log_w: np.ndarray, u0: int) -> np.ndarray:
    """Ancestor indices (int64[N]) for the offset u0 in [0, 2^32)."""
    q = fixed_point_weights(log_w)
    n = len(q)
    c = np.cumsum(q, dtype=np.uint64)
    total = int(c[-1])
    assert total > 0, "all weights are zero"
    c_scaled = [int(v) * (n << 32) for v in c]
    out = np.empty(n, dtype=np.int64)
    for k in range(n):
        thr = ((k << 32) + int(u0)) * total
        out[k] = bisect.bisect_right(c_scaled, thr)
    return out


def ess_from_log_w(log_w: np.ndarray) -> float:
    lw = np.asarray(log_w, dtype=np.float64)
    w = np.exp(lw - lw.max())
    return float(w.sum() ** 2 / (w ** 2).sum() / len(lw))
