"""Oracle restatement of the reference AIS sampler and its transition operators.
TEST INFRASTRUCTURE (CPU, plain PyTorch).  Pinned bit-for-bit against the reference by
`oracle/gen_golden.py` (same seeds => same numbers on CPU).

Follows, statement by statement in its own words:
  fab/sampling_methods/base.py:7-118        Point, value+grad, gamma, grad-gamma
  fab/sampling_methods/ais.py:20-213        AnnealedImportanceSampler
  fab/sampling_methods/transition_operators/hmc.py:8-202       HMC
  fab/sampling_methods/transition_operators/metropolis.py:9-74 Metropolis
  fab/utils/numerical.py:18-23              effective_sample_size
Every reference quirk listed in SURVEY Appendix A.3 is reproduced on purpose.
The only addition is the `noise` hook (see oracle/noise.py), which by default issues
the reference's own torch RNG calls in the reference's order.
"""
import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from oracle.noise import TorchNoise


# ----------------------------------------------------------------------------- Point
class Point:
    """SoA record of chain state (base.py:7-47)."""

    def __init__(self, x, log_q, log_p, grad_log_q=None, grad_log_p=None):
        self.x, self.log_q, self.log_p = x, log_q, log_p
        self.grad_log_q, self.grad_log_p = grad_log_q, grad_log_p

    @property
    def device(self):
        return self.x.device

    def __getitem__(self, idx):
        gq = None if self.grad_log_q is None else self.grad_log_q[idx]
        gp = None if self.grad_log_p is None else self.grad_log_p[idx]
        return Point(self.x[idx], self.log_q[idx], self.log_p[idx], gq, gp)

    def __setitem__(self, idx, other):
        self.x[idx] = other.x
        self.log_q[idx] = other.log_q
        self.log_p[idx] = other.log_p
        if self.grad_log_q is not None:
            self.grad_log_q[idx] = other.grad_log_q
            self.grad_log_p[idx] = other.grad_log_p


def value_and_input_grad(x, fn):
    """base.py:50-56: autograd value + d/dx with a ones cotangent."""
    x = x.detach()
    x.requires_grad = True
    y = fn(x)
    g = torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), retain_graph=True)[0]
    return g.detach(), y.detach()


def make_point(x, log_q_fn, log_p_fn, with_grad: bool, log_q_x=None) -> Point:
    """base.py:59-72.  NB with_grad=True ignores a supplied log_q_x (quirk 7)."""
    x = x.detach()
    if with_grad:
        gq, lq = value_and_input_grad(x, log_q_fn)
        gp, lp = value_and_input_grad(x, log_p_fn)
        return Point(x=x, log_q=lq, log_p=lp, grad_log_q=gq, grad_log_p=gp)
    lq = log_q_x if log_q_x is not None else log_q_fn(x)
    return Point(x=x, log_q=lq.detach(), log_p=log_p_fn(x).detach())


def gamma(pt: Point, beta, alpha, p_target: bool):
    """Interpolated log-density (base.py:76-97)."""
    with torch.no_grad():
        if p_target:
            return (1 - beta) * pt.log_q + beta * pt.log_p
        return ((1 - beta) + beta * (1 - alpha)) * pt.log_q + beta * alpha * pt.log_p


def grad_gamma(pt: Point, beta, alpha, p_target: bool):
    """base.py:100-118.  The literal 2 (not alpha) in front of grad_log_p is quirk 1."""
    with torch.no_grad():
        if p_target:
            return (1 - beta) * pt.grad_log_q + beta * pt.grad_log_p
        return ((1 - beta) + beta * (1 - alpha)) * pt.grad_log_q + 2 * beta * pt.grad_log_p


def effective_sample_size(log_w: torch.Tensor) -> torch.Tensor:
    """numerical.py:18-23."""
    assert log_w.dim() == 1
    w = F.softmax(log_w, dim=0)
    return 1 / torch.sum(w ** 2) / log_w.shape[0]


def beta_schedule(kind: str, n: int) -> torch.Tensor:
    """ais.py:108-129; float64 tensor of length n+2."""
    assert n > 0
    if kind == "linear":
        b = np.linspace(0.0, 1.0, n + 2)
    elif kind == "geometric":
        n_lin = int(n / 4)
        n_geo = n - n_lin - 1
        b = np.concatenate([np.linspace(0, 0.01, n_lin + 2)[:-1],
                            np.geomspace(0.01, 1, n_geo + 2)])
    else:
        raise Exception(f"distribution spacing incorrectly specified: '{kind}',"
                        f"options are 'geometric' or 'linear'")
    assert b.shape == (n + 2,)
    return torch.tensor(b)


# ------------------------------------------------------------------- transition operators
class _Operator(torch.nn.Module):
    """transition_operators/base.py:12-85."""
    uses_grad_info = False

    def __init__(self, n_dist, dim, base_log_prob, target_log_prob, p_target, alpha):
        self.dim = dim
        self.target_log_prob = target_log_prob
        self.base_log_prob = base_log_prob
        self.alpha = alpha
        self.n_ais_intermediate_distributions = n_dist
        self.p_target = p_target
        self.noise = TorchNoise()
        super().__init__()

    def new_point(self, x) -> Point:
        return make_point(x, self.base_log_prob, self.target_log_prob,
                          with_grad=self.uses_grad_info)

    def g(self, pt, beta):
        return gamma(pt, beta, self.alpha, self.p_target)

    def dg(self, pt, beta):
        return grad_gamma(pt, beta, self.alpha, self.p_target)


class OracleHMC(_Operator):
    """hmc.py:8-202 (defaults as in the reference ctor, hmc.py:9-24)."""
    uses_grad_info = True

    def __init__(self, n_ais_intermediate_distributions, dim, base_log_prob, target_log_prob,
                 alpha=None, p_target=False, epsilon=1.0, n_outer=1, L=5, mass_init=1.0,
                 target_p_accept=0.65, max_grad=1e3, common_epsilon_init_weight=0.1,
                 eval_mode=False):
        super().__init__(n_ais_intermediate_distributions, dim, base_log_prob, target_log_prob,
                         p_target, alpha)
        w = common_epsilon_init_weight
        self.register_buffer("common_epsilon", torch.tensor([epsilon * w]))
        self.register_buffer("epsilons",
                             torch.ones([n_ais_intermediate_distributions, n_outer])
                             * epsilon * (1 - w))
        self.register_buffer("mass_vector", torch.ones(dim) * mass_init)
        self.n_outer, self.L = n_outer, L
        self.target_p_accept, self.max_grad = target_p_accept, max_grad
        self.eval_mode = eval_mode
        self.first_dist_p_accepts = [torch.tensor([0.0]) for _ in range(n_outer)]
        self.last_dist_p_accepts = [torch.tensor([0.0]) for _ in range(n_outer)]

    def set_eval_mode(self, flag: bool):
        self.eval_mode = flag

    def step_size(self, i: int, n: int):
        return self.epsilons[i - 1, n] + self.common_epsilon  # hmc.py:90-100

    def get_logging_info(self) -> dict:
        M = self.n_ais_intermediate_distributions
        d = {}
        for k, v in enumerate(self.first_dist_p_accepts):
            d[f"dist0_p_accept_{k}"] = v.item()
        if M > 1:
            for k, v in enumerate(self.last_dist_p_accepts):
                d[f"dist{M - 1}_p_accept_{k}"] = v.item()
        d["epsilons_dist0_loop0"] = self.step_size(0, 0).cpu().item()
        if M > 1:
            d[f"epsilons_dist{M - 1}_loop0"] = self.step_size(M - 1, 0).cpu().item()
        d["average_distance_dist0"] = self.average_distance_first_dist.cpu().item()
        if hasattr(self, "average_distance_last_dist"):
            d[f"average_distance_dist_{M - 1}"] = self.average_distance_last_dist.cpu().item()
        return d

    def _grad_U(self, pt, beta):
        g = -self.dg(pt, beta)
        return torch.nan_to_num(torch.clamp(g, max=self.max_grad, min=-self.max_grad),
                                nan=0.0, posinf=0.0, neginf=0.0)       # hmc.py:194-199

    def _log_joint(self, pt, p, beta):
        return self.g(pt, beta) - torch.sum(p ** 2 / self.mass_vector, dim=-1) / 2

    def _accept(self, prop, cur, p_prop, p_cur, beta):
        lp_cur = self._log_joint(cur, p_cur, beta)
        lp_prop = self._log_joint(prop, p_prop, beta)
        with torch.no_grad():
            log_a = lp_prop - lp_cur
            ok = torch.isfinite(log_a)
            ninf = -float("inf")
            log_a = torch.nan_to_num(log_a, nan=ninf, posinf=ninf, neginf=ninf)
            e = self.noise.exponential(log_a.shape, log_a.device)
            accept = (log_a > -e) & ok
            log_a = torch.clamp(log_a, max=0.0)
            log_mean = torch.logsumexp(log_a, dim=-1) - \
                torch.log(torch.tensor(log_a.shape[0])).to(log_a.device)
            return accept, log_mean

    def _tune(self, log_mean, i, n):
        thr = torch.log(torch.tensor(self.target_p_accept).to(log_mean.device))
        if log_mean > thr:
            self.epsilons[i - 1, n] = self.epsilons[i - 1, n] * 1.05
            self.common_epsilon = self.common_epsilon * 1.02
        else:
            self.epsilons[i - 1, n] = self.epsilons[i - 1, n] / 1.05
            self.common_epsilon = self.common_epsilon / 1.02

    def _log_stats(self, i, n, p_mean, moved_x, origin_x):
        if i == 1:
            self.first_dist_p_accepts[n] = p_mean.cpu().detach()
            dist = torch.linalg.norm(origin_x - moved_x, ord=2, dim=-1)
            self.average_distance_first_dist = torch.mean(dist).detach().cpu()
        elif i == self.n_ais_intermediate_distributions:
            self.last_dist_p_accepts[n] = p_mean.cpu().detach()
            dist = torch.linalg.norm(origin_x - moved_x, ord=2, dim=-1)
            self.average_distance_last_dist = torch.mean(dist).detach().cpu()

    def transition(self, point: Point, i: int, beta) -> Point:
        cur = point                          # alias, mutated in place (hmc.py:130,154)
        for n in range(self.n_outer):
            origin = cur                     # alias of cur (so "distance" is post-accept, quirk 3)
            eps = self.step_size(i, n)
            p = self.noise.momentum(point.x) * self.mass_vector
            p0 = p
            gu = self._grad_U(point, beta)   # at the previous *proposal* for n>0 (quirk 2)
            for _ in range(self.L):
                p = p - eps * gu / 2
                x = point.x + eps / self.mass_vector * p
                point = self.new_point(x)
                gu = self._grad_U(point, beta)
                p = p - eps * gu / 2
            accept, log_mean = self._accept(point, cur, p, p0, beta)
            cur[accept] = point[accept]
            self._log_stats(i, n, torch.exp(log_mean), point.x, origin.x)
            if not self.eval_mode:
                self._tune(log_mean, i, n)
        return cur


class OracleMetropolis(_Operator):
    """metropolis.py:9-74."""
    uses_grad_info = False

    def __init__(self, n_ais_intermediate_distributions, dim, base_log_prob, target_log_prob,
                 n_updates, alpha=None, p_target=False, max_step_size=1.0, min_step_size=0.1,
                 adjust_step_size=True, target_p_accept=0.65, eval_mode=False):
        super().__init__(n_ais_intermediate_distributions, dim, base_log_prob, target_log_prob,
                         p_target, alpha)
        self.n_updates = n_updates
        self.adjust_step_size = adjust_step_size
        self.register_buffer(
            "noise_scalings",
            torch.linspace(max_step_size, min_step_size, n_updates).repeat(
                (n_ais_intermediate_distributions, 1)))
        self.target_prob_accept = target_p_accept
        self.eval_mode = eval_mode

    def set_eval_mode(self, flag: bool):
        self.eval_mode = not flag            # inverted in the reference (quirk 4)

    def get_logging_info(self) -> Dict:
        return {"noise_scaling_0_0": self.noise_scalings[0, 0].cpu().item(),
                "noise_scaling_0_-1": self.noise_scalings[0, -1].cpu().item()}

    def transition(self, point: Point, i: int, beta) -> Point:
        g_prev = self.g(point, beta)         # never refreshed after an accept (quirk 5)
        for n in range(self.n_updates):
            x = point.x
            x_new = x + self.noise.proposal(x) * self.noise_scalings[i - 1, n]
            prop = self.new_point(x_new)
            a = torch.exp(self.g(prop, beta) - g_prev)
            a = torch.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0)
            accept = (a > self.noise.uniform(a.shape, x.device)).int()
            point[accept.bool()] = prop[accept.bool()]
            if self.adjust_step_size and not self.eval_mode:
                p_acc = torch.mean(torch.clamp_max(a, 1))
                if p_acc > self.target_prob_accept:
                    self.noise_scalings[i - 1, n] = self.noise_scalings[i - 1, n] * 1.05
                else:
                    self.noise_scalings[i - 1, n] = self.noise_scalings[i - 1, n] / 1.05
        return point


# ------------------------------------------------------------------------------- AIS
class OracleAIS:
    """ais.py:20-213."""

    def __init__(self, base_distribution, target_log_prob: Callable, transition_operator,
                 p_target: bool, alpha: Optional[float] = None,
                 n_intermediate_distributions: int = 1,
                 distribution_spacing_type: str = "linear"):
        if not p_target:
            assert alpha is not None, "Must specify alpha if AIS target is not p."
        self.base_distribution = base_distribution
        self.target_log_prob = target_log_prob
        self.transition_operator = transition_operator
        self.p_target, self.alpha = p_target, alpha
        self.n_intermediate_distributions = n_intermediate_distributions
        self.B_space = beta_schedule(distribution_spacing_type, n_intermediate_distributions)
        self._logging_info = None

    def get_logging_info(self):
        d = dict(self._logging_info)
        d.update(self.transition_operator.get_logging_info())
        return d

    def _drop_non_finite(self, pt: Point, log_w, where: str, raise_exception=True):
        ok = ~torch.isinf(pt.log_p) & ~torch.isnan(pt.log_p) & \
             ~torch.isinf(pt.log_q) & ~torch.isnan(pt.log_q)
        if torch.sum(ok) == 0:
            if raise_exception:
                raise Exception(f"No valid points generated in sampling the {where}")
            return pt, log_w
        return pt[ok], log_w[ok]

    def _step(self, pt: Point, log_w, j: int):
        pt = self.transition_operator.transition(pt, j, self.B_space[j])
        if self.B_space[j + 1] != self.B_space[j]:
            num = gamma(pt, self.B_space[j + 1], self.alpha, self.p_target)
            den = gamma(pt, self.B_space[j], self.alpha, self.p_target)
            log_w = log_w + (num - den)
        return pt, log_w

    def sample_and_log_weights(self, batch_size: int, logging: bool = True
                               ) -> Tuple[Point, torch.Tensor]:
        op = self.transition_operator
        x, log_q0 = self.base_distribution.sample_and_log_prob((batch_size,))
        pt = make_point(x, self.base_distribution.log_prob, self.target_log_prob,
                        with_grad=op.uses_grad_info, log_q_x=log_q0)
        log_w = gamma(pt, self.B_space[1], self.alpha, self.p_target) - log_q0
        pt, log_w = self._drop_non_finite(pt, log_w, "chain init")
        if logging:
            with torch.no_grad():
                ess_base = effective_sample_size(pt.log_p - pt.log_q).detach().cpu().item()
        for j in range(1, self.n_intermediate_distributions + 1):
            pt, log_w = self._step(pt, log_w, j)
        pt, log_w = self._drop_non_finite(pt, log_w, "chain end")
        if logging:
            with torch.no_grad():
                ess_ais = effective_sample_size(log_w).cpu().item()
                lse = torch.logsumexp(log_w, dim=0)
                log_Z = lse - torch.log(torch.ones_like(lse) * batch_size)   # quirk 6
                self._logging_info = dict(ess_base=ess_base, ess_ais=ess_ais,
                                          log_Z=log_Z.cpu().item())
        return pt, log_w.detach()

    def generate_eval_data(self, outer_batch_size: int, inner_batch_size: int):
        """ais.py:132-188."""
        assert outer_batch_size % inner_batch_size == 0
        op = self.transition_operator
        out = ([], [], [], [])
        for _ in range(outer_batch_size // inner_batch_size):
            x, log_q0 = self.base_distribution.sample_and_log_prob((inner_batch_size,))
            pt = make_point(x, self.base_distribution.log_prob, self.target_log_prob,
                            with_grad=op.uses_grad_info, log_q_x=log_q0)
            base_log_w = self.target_log_prob(x) - log_q0
            pt, base_log_w = self._drop_non_finite(pt, base_log_w, "chain init")
            out[0].append(pt.x.detach().cpu())
            out[1].append(base_log_w.detach().cpu())
            log_w = gamma(pt, self.B_space[1], self.alpha, self.p_target) - pt.log_q
            for j in range(1, self.n_intermediate_distributions + 1):
                pt, log_w = self._step(pt, log_w, j)
            pt, log_w = self._drop_non_finite(pt, log_w, "chain end", raise_exception=False)
            out[2].append(pt.x.detach().cpu())
            out[3].append(log_w.detach().cpu())
        return tuple(torch.cat(o, dim=0) for o in out)
