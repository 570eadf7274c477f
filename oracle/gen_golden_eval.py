"""Pin oracle/eval_manywell.py against the UNMODIFIED reference and write
tests/golden/eval_manywell.pt.  TEST INFRASTRUCTURE (build container only):

    python -m oracle.gen_golden_eval

Checks (all bit-for-bit on CPU, identical torch seeds on both sides):
  * the mode test set for d = 2, 4, 8, 32 (many_well.py:27-35);
  * the exact sampler (many_well.py:61-67, double_well.py:60-94, rejection_sampling.py:6-20);
  * `performance_metrics` without and with a `log_q_fn`, batch sizes 10 and 64 (many_well.py:96-147).
The fixture holds the seeded reference outputs for d = 4 and d = 8.
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import eval_manywell as ev                      # noqa: E402
from oracle.targets import OracleManyWell                   # noqa: E402
from oracle.ref_loader import load_reference                # noqa: E402


def std_normal_log_prob(x):
    return -0.5 * (x ** 2).sum(-1) - 0.5 * x.shape[-1] * math.log(2 * math.pi)


def main():
    load_reference()
    from fab.target_distributions.many_well import ManyWellEnergy as Ref
    checks, fixture = 0, {}
    for dim in (2, 4, 8, 32):
        ref = Ref(dim, use_gpu=False)
        assert torch.equal(ref._test_set_modes, ev.mode_test_set(dim)), f"mode set d={dim}"
        checks += 1
    for dim in (4, 8):
        ref, orc = Ref(dim, use_gpu=False), OracleManyWell(dim)
        torch.manual_seed(100 + dim)
        xs_ref = ref.sample((257,))
        torch.manual_seed(100 + dim)
        xs_orc = ev.sample_many_well(dim, (257,))
        assert torch.equal(xs_ref, xs_orc), f"sampler d={dim}"
        checks += 1
        g = torch.Generator().manual_seed(dim)
        log_w = torch.randn(5025, generator=g) * 1.5 + float(orc.log_Z)
        m_ref = ref.performance_metrics(None, log_w)
        m_orc = ev.performance_metrics(orc, log_w)
        assert m_ref == m_orc, (m_ref, m_orc)
        checks += 1
        entry = dict(seed=100 + dim, samples=xs_ref, log_w=log_w, metrics_no_q=m_ref, metrics_q={})
        for bs in (10, 64):
            torch.manual_seed(7 * dim + bs)
            a = ref.performance_metrics(None, log_w, std_normal_log_prob, bs)
            torch.manual_seed(7 * dim + bs)
            b = ev.performance_metrics(orc, log_w, std_normal_log_prob, bs)
            assert a == b, (a, b)
            checks += 1
            entry["metrics_q"][bs] = dict(seed=7 * dim + bs, info=a)
        # chunk sizes the reference's iterator yields
        it = ref.get_modes_test_set_iterator(10)
        sizes_ref = [x.shape[0] for x in it]
        it2, n = ev.modes_test_set(dim, 10)
        assert sizes_ref == [x.shape[0] for x in it2] and n == it.test_set_n_points
        checks += 1
        entry["iterator_chunk_sizes_bs10"] = sizes_ref
        fixture[dim] = entry
    out = os.path.join(ROOT, "tests", "golden", "eval_manywell.pt")
    torch.save(fixture, out)
    print(f"[gen_golden_eval] {checks} checks against the reference passed; wrote {out}")
    gmm_checks()


def gmm_checks():
    """GMM target (gmm.py:71-99, numerical.py:8-60): quadratic test function, importance-weighted
    bias metrics, and the model-density branch with callables standing in for the densities."""
    from fab.target_distributions.gmm import GMM as Ref
    from fab.utils.numerical import quadratic_function as ref_quad
    from oracle import eval_gmm as eg
    checks, fixture = 0, {}
    for dim, n_mixes, loc in ((2, 4, 8.0), (3, 7, 5.0)):
        torch.manual_seed(dim)
        ref = Ref(dim=dim, n_mixes=n_mixes, loc_scaling=loc, log_var_scaling=1.0, use_gpu=False,
                  true_expectation_estimation_n_samples=int(1e5))
        g = torch.Generator().manual_seed(50 + dim)
        x = torch.randn(3000, dim, generator=g) * loc
        log_w = torch.randn(3000, generator=g) * 2
        assert torch.equal(ref_quad(x), eg.quadratic_function(x)), "quadratic function"
        checks += 1
        m_ref = ref.performance_metrics(x, log_w)
        assert m_ref == eg.performance_metrics(x, log_w, ref.true_expectation), "metrics without q"
        checks += 1
        log_q = lambda z: std_normal_log_prob(z / loc) - z.shape[-1] * math.log(loc)
        torch.manual_seed(900 + dim)
        m_q_ref = ref.performance_metrics(x, log_w, log_q)
        torch.manual_seed(900 + dim)
        m_q = eg.performance_metrics(x, log_w, ref.true_expectation, log_q, ref.log_prob,
                                     lambda: ref.sample((ref.n_test_set_samples,)))
        # the reference re-seeds the global RNG inside quadratic_function (`torch.seed()`,
        # numerical.py:40), so its two test sets cannot be reproduced by seeding: the bias terms
        # must agree exactly, the test-set statistics (1000 fresh samples each) only statistically
        for k in ("bias_normed", "bias_no_correction"):
            assert m_q_ref[k] == m_q[k], (k, m_q_ref, m_q)
        for k, tol in (("test_set_mean_log_prob", 0.2), ("kl_forward", 0.2)):
            assert abs(m_q_ref[k] - m_q[k]) < tol, (k, m_q_ref, m_q)
        assert 0.5 < m_q_ref["ess_over_p"] / m_q["ess_over_p"] < 2.0 and set(m_q_ref) == set(m_q)
        checks += 1
        fixture[dim] = dict(ctor_seed=dim, n_mixes=n_mixes, loc_scaling=loc, locs=ref.locs.clone(),
                            true_expectation=ref.true_expectation.clone(), x=x, log_w=log_w,
                            quad=ref_quad(x), metrics_no_q=m_ref, metrics_q=m_q_ref, q_seed=900 + dim)
    out = os.path.join(ROOT, "tests", "golden", "eval_gmm.pt")
    torch.save(fixture, out)
    print(f"[gen_golden_eval] GMM: {checks} checks against the reference passed; wrote {out}")


if __name__ == "__main__":
    main()
