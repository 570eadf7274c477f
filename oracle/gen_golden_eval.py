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


if __name__ == "__main__":
    main()
