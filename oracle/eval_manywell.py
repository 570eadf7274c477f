"""Oracle restatement of the Many-Well evaluation path.  TEST INFRASTRUCTURE.

  fab/sampling_methods/rejection_sampling.py:6-20      rejection_sampling
  fab/target_distributions/double_well.py:60-94        DoubleWellEnergy.sample (+ first dimension)
  fab/target_distributions/many_well.py:26-35,61-79    mode test set, sample, test-set iterator
  fab/target_distributions/many_well.py:96-147         performance_metrics
  fab/utils/training.py:36-53                          DatasetIterator

CPU only; pinned bit-for-bit against the unmodified reference by `python -m oracle.gen_golden_eval`
(seeded sampler draws, the mode test set, and performance_metrics with and without a log_q_fn),
which also writes tests/golden/eval_manywell.pt.  Only tests may import this module.

Reference quirks kept (they change numbers):
  * `log_w.split(50)` makes chunks of SIZE 50, stacked on the last axis -> 50 log-Z estimates of
    n/50 weights each (many_well.py:101-104);
  * the number of exact-sample batches is `max(50 // batch_size, 1)` because it is read off the
    stacked tensor's first axis (many_well.py:119);
  * DatasetIterator yields chunks of size ceil(N / batch_size), not of size batch_size
    (training.py:41-42).
"""
import math
from typing import Callable, Dict, Iterator, Optional

import numpy as np
import torch

from oracle.targets import OracleManyWell, DOUBLE_WELL_Z

CENTRE = 1.7                      # many_well.py:25
MAX_DIM_FOR_ALL_MODES = 40        # many_well.py:26


def first_dim_log_target(x: torch.Tensor) -> torch.Tensor:
    """double_well.py:65-66: -x^4 + 6 x^2 + x/2 (same operation order)."""
    t = -(x ** 4)
    t = t + 6 * x ** 2
    return t + 1 / 2 * x


def proposal(device="cpu"):
    """double_well.py:38-41,71-75: 0.2 N(-1.7, 0.5) + 0.8 N(1.7, 0.5)."""
    mix = torch.distributions.Categorical(torch.tensor([0.2, 0.8], device=device))
    com = torch.distributions.Normal(torch.tensor([-1.7, 1.7], device=device),
                                     torch.tensor([0.5, 0.5], device=device))
    return torch.distributions.MixtureSameFamily(mixture_distribution=mix, component_distribution=com)


def rejection_sample_first_dim(n: int, device="cpu") -> torch.Tensor:
    """rejection_sampling.py:6-20 with k = 3 Z (double_well.py:77); the recursion of the reference
    is a loop here (same RNG calls in the same order)."""
    k = DOUBLE_WELL_Z * 3
    prop = proposal(device)
    parts, need = [], n
    while need > 0:
        z = prop.sample((need * 10,))
        u = torch.distributions.Uniform(0, k * torch.exp(prop.log_prob(z))).sample().to(z)
        got = z[torch.exp(first_dim_log_target(z)) > u][:need]
        parts.append(got)
        need -= got.shape[0]
    return torch.concat(parts, dim=0) if len(parts) > 1 else parts[0]


def sample_many_well(dim: int, shape, device="cpu") -> torch.Tensor:
    """many_well.py:61-67 + double_well.py:84-92: per well, first coordinate by rejection sampling
    then second ~ N(0, 1)."""
    assert len(shape) == 1
    wells = []
    for _ in range(dim // 2):
        x1 = rejection_sample_first_dim(shape[0], device)
        x2 = torch.distributions.Normal(torch.tensor(0.0).to(x1.device),
                                        torch.tensor(1.0).to(x1.device)).sample(shape)
        wells.append(torch.stack([x1, x2], dim=-1))
    return torch.concat(wells, dim=-1)


def mode_test_set(dim: int) -> torch.Tensor:
    """many_well.py:27-35: every sign pattern of +-1.7 on the first coordinate of each well (first
    well slowest), zeros elsewhere."""
    n = dim // 2
    rows = torch.arange(2 ** n)
    bits = (rows[:, None] >> torch.arange(n - 1, -1, -1)[None, :]) & 1
    out = torch.zeros((2 ** n, dim))
    out[:, 0::2] = (2.0 * bits - 1.0) * CENTRE
    return out


def chunk_iterator(dataset: torch.Tensor, batch_size: int):
    """training.py:36-53.  Returns (iterator, n_points)."""
    n_splits = int(np.ceil(dataset.shape[0] / batch_size))
    return iter(torch.split(dataset, n_splits)), dataset.shape[0]


def modes_test_set(dim: int, batch_size: int):
    """many_well.py:69-79."""
    if dim < MAX_DIM_FOR_ALL_MODES:
        data = mode_test_set(dim)
    else:
        n = int(1e4)
        data = torch.zeros((n, dim))
        data[:, torch.arange(dim) % 2 == 0] = \
            -CENTRE + CENTRE * 2 * torch.randint(high=2, size=(n, int(dim / 2)))
    return chunk_iterator(data, batch_size)


def performance_metrics(target: OracleManyWell, log_w: torch.Tensor,
                        log_q_fn: Optional[Callable] = None, batch_size: Optional[int] = None,
                        sampler: Optional[Callable] = None) -> Dict:
    """many_well.py:96-147.  `sampler(shape)` defaults to the exact sampler above."""
    n_runs = 50
    keep = (log_w.shape[0] // n_runs) * n_runs
    stacked = torch.stack(log_w[:keep].split(n_runs), dim=-1)
    log_Z_estimate = torch.logsumexp(stacked, dim=-1) - np.log(stacked.shape[-1])
    relative_error = torch.exp(log_Z_estimate - target.log_Z) - 1
    info = dict(relative_MSE_Z_estimate=torch.mean(torch.abs(relative_error)).cpu().item(),
                abs_MSE_log_Z_estimate=torch.mean(torch.abs(log_Z_estimate - target.log_Z)).cpu().item())
    if log_q_fn is None:
        return info
    assert batch_size is not None
    n_batches = max(stacked.shape[0] // batch_size, 1)
    sampler = sampler or (lambda shape: sample_many_well(target.dim, shape))
    s_modes, s_exact, s_kl = 0.0, 0.0, 0.0
    it, n_points = modes_test_set(target.dim, batch_size)
    for x in it:
        s_modes += torch.sum(log_q_fn(x)).detach().cpu()
    for _ in range(n_batches):
        x = sampler((batch_size,))
        lq = log_q_fn(x)
        s_exact += torch.sum(lq).detach().cpu()
        s_kl += torch.sum(target.log_prob(x) - target.log_Z - lq).detach().cpu()
    m = batch_size * n_batches
    info.update(test_set_modes_mean_log_prob=(s_modes / n_points).cpu().item(),
                test_set_exact_mean_log_prob=(s_exact / m).cpu().item(),
                forward_kl=(s_kl / m).cpu().item(), eval_batch_size=m)
    return info
