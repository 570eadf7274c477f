"""Oracle restatement of the GMM target's evaluation metrics.  TEST INFRASTRUCTURE (CPU only).

  fab/utils/numerical.py:8-15,25-30,33-60   MC estimate, ESS over p, quadratic test function,
                                            importance-weighted expectation
  fab/target_distributions/gmm.py:53-55,71-99   test_set, evaluate_expectation, performance_metrics

Pinned bit-for-bit against the unmodified reference by `python -m oracle.gen_golden_eval`
(fixture tests/golden/eval_gmm.pt).  Differences on purpose: the quadratic function's tables are
drawn from a LOCAL generator seeded like the reference seeds the global one (same values, but the
reference's `torch.seed()` re-seeding of the global RNG on every call is not reproduced).
Consequence: the reference's model-density branch draws its test sets from an unpredictable seed, so
that branch is pinned exactly only in its deterministic part (the two bias terms) and statistically
in the rest.  Reference quirk kept: `test_set` is re-sampled on every access, so log q and log p of the
model-density branch are evaluated on two different sample sets (gmm.py:84-85).
"""
from typing import Callable, Dict, Optional

import torch


def quadratic_tables(dim: int, seed: int = 0):
    """numerical.py:33-45: x_shift = 2 randn(d), A = 2 rand(d, d), b = rand(d) after manual_seed."""
    g = torch.Generator().manual_seed(seed)
    x_shift = 2 * torch.randn(dim, generator=g)
    A = 2 * torch.rand((dim, dim), generator=g)
    b = torch.rand(dim, generator=g)
    return x_shift, A, b


def quadratic_function(x: torch.Tensor, seed: int = 0) -> torch.Tensor:
    """numerical.py:48-51."""
    x_shift, A, b = (t.to(x) for t in quadratic_tables(x.shape[-1], seed))
    x = x + x_shift
    return torch.einsum("bi,ij,bj->b", x, A, x) + torch.einsum("i,bi->b", b, x)


def evaluate_expectation(samples, log_w, true_expectation) -> torch.Tensor:
    """gmm.py:71-77 + numerical.py:55-60: normalised bias of the importance-weighted estimate."""
    w = torch.softmax(log_w, dim=-1)
    est = w @ quadratic_function(samples)       # (the reference writes w.T @ f: same dot product)
    true_expectation = true_expectation.to(est.device)
    return (est - true_expectation) / true_expectation


def performance_metrics(samples, log_w, true_expectation, log_q_fn: Optional[Callable] = None,
                        log_p_fn: Optional[Callable] = None,
                        test_set_fn: Optional[Callable] = None) -> Dict:
    """gmm.py:79-99.  `test_set_fn()` draws a fresh test set (called twice, like the property)."""
    bias_normed = evaluate_expectation(samples, log_w, true_expectation)
    bias_no_correction = evaluate_expectation(samples, torch.ones_like(log_w), true_expectation)
    if not log_q_fn:
        return {"bias_normed": bias_normed.cpu().item(),
                "bias_no_correction": torch.abs(bias_no_correction).cpu().item()}
    log_q_test = log_q_fn(test_set_fn())
    log_p_test = log_p_fn(test_set_fn())
    return {"test_set_mean_log_prob": torch.mean(log_q_test).cpu().item(),
            "bias_normed": torch.abs(bias_normed).cpu().item(),
            "bias_no_correction": torch.abs(bias_no_correction).cpu().item(),
            "ess_over_p": (1 / torch.mean(torch.exp(log_p_test - log_q_test))).detach().cpu().item(),
            "kl_forward": torch.mean(log_p_test - log_q_test).detach().cpu().item()}
