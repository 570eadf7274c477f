"""Transition operators behind the reference's `TransitionOperator` surface
(fab/sampling_methods/transition_operators/base.py:12-85), each transition one fused launch.

  HamiltonianMonteCarlo  fab/sampling_methods/transition_operators/hmc.py:8-202
  Metropolis             fab/sampling_methods/transition_operators/metropolis.py:9-74

Constructor signatures, registered buffer names/shapes (`common_epsilon[1]`, `epsilons[M,n_outer]`,
`mass_vector[d]`, `noise_scalings[M,n_updates]` -- checkpointed by fab/core.py:226-228), logging
keys and every quirk of SURVEY Appendix A.3 are kept.  The step-size tuner and the logging scalars
live on the device, so a transition never synchronises with the host.

The operators need to know *what* `base_log_prob` / `target_log_prob` compute in order to fuse
them: both must be bound methods of objects from this package (`B200RealNVP.log_prob`,
`ManyWellEnergy.log_prob`, ...).  Anything else raises -- there is no slow path.
"""
import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from fab_torch_b200 import _lib
from fab_torch_b200 import dist as fdist
from fab_torch_b200.point import Point
from fab_torch_b200.types_ import LogProbFunc


def make_gamma(beta, alpha, p_target: bool) -> "_lib.Gamma":
    """Coefficients of gamma / grad-gamma (fab/sampling_methods/base.py:94,97,116,118), computed in
    float64 like the reference (beta is a 0-dim float64 tensor there) and rounded to fp32 at the
    point where torch would multiply them into an fp32 tensor."""
    b = float(beta)
    if p_target:
        cq, cp, gq, gp = 1 - b, b, 1 - b, b
    else:
        assert alpha is not None, "Must specify alpha if AIS target is not p."
        a = float(alpha)
        cq = (1 - b) + b * (1 - a)
        cp = b * a
        gq = cq
        gp = 2 * b                     # literal 2, not alpha (quirk 1)
    return _lib.Gamma(cq, cp, gq, gp)


class DeviceNoise:
    """Default randomness: torch's CUDA generator, all draws of one transition in one call."""

    def base_eps(self, n, d, device):
        return torch.randn((n, d), dtype=torch.float32, device=device)

    def momentum(self, i, n_outer, n, d, device):
        return torch.randn((n_outer, n, d), dtype=torch.float32, device=device)

    def exponential(self, i, n_outer, n, device):
        return torch.empty((n_outer, n), dtype=torch.float32, device=device).exponential_(1.0)

    def proposal(self, i, n_updates, n, d, device):
        return torch.randn((n_updates, n, d), dtype=torch.float32, device=device)

    def uniform(self, i, n_updates, n, device):
        return torch.rand((n_updates, n), dtype=torch.float32, device=device)


class InjectedNoise(DeviceNoise):
    """Replays pre-drawn noise (parity tests, reproducible runs).  `record` maps the keys
    base_eps / momentum / exponential / proposal / uniform to lists of tensors in the order the
    reference would have drawn them (oracle/noise.py)."""

    def __init__(self, record: Dict):
        self._it = {k: iter(v) for k, v in record.items()}

    def _take(self, key, count, device):
        ts = [next(self._it[key]).to(device=device, dtype=torch.float32) for _ in range(count)]
        return torch.stack(ts).contiguous()

    def base_eps(self, n, d, device):
        return self._take("base_eps", 1, device)[0]

    def momentum(self, i, n_outer, n, d, device):
        return self._take("momentum", n_outer, device)

    def exponential(self, i, n_outer, n, device):
        return self._take("exponential", n_outer, device)

    def proposal(self, i, n_updates, n, d, device):
        return self._take("proposal", n_updates, device)

    def uniform(self, i, n_updates, n, device):
        return self._take("uniform", n_updates, device)


def _owner(fn, what):
    obj = getattr(fn, "__self__", None)
    if obj is None:
        raise TypeError(f"{what} must be a bound method of a fab_torch_b200 object "
                        f"(got {fn!r}); the fused kernels cannot call arbitrary Python")
    return obj


class TransitionOperator(nn.Module):
    def __init__(self, n_ais_intermediate_distributions: int, dim: int,
                 base_log_prob: LogProbFunc, target_log_prob: LogProbFunc,
                 p_target: bool = True, alpha: float = None):
        self.dim = dim
        self.target_log_prob = target_log_prob
        self.base_log_prob = base_log_prob
        self.alpha = alpha
        self.n_ais_intermediate_distributions = n_ais_intermediate_distributions
        self.p_target = p_target
        super().__init__()
        flow = _owner(base_log_prob, "base_log_prob")
        target = _owner(target_log_prob, "target_log_prob")
        if not (hasattr(flow, "desc") and hasattr(flow, "blob")):
            raise TypeError("base_log_prob must be B200RealNVP.log_prob")
        if not hasattr(target, "target_desc"):
            raise TypeError("target_log_prob must be the log_prob of a fab_torch_b200 target")
        # plain attributes (not submodules): the flow is owned by the caller
        object.__setattr__(self, "_flow", flow)
        object.__setattr__(self, "_target", target)
        self.noise = DeviceNoise()
        self.chain_noise_override = None  # one-shot pre-drawn noise for the next chain (e.g. from host)
        self.process_group = None       # set by the sampler for particle-parallel runs
        self.register_buffer("_stats", torch.zeros(2 * _lib.FAB_MAX_UPDATES), persistent=False)
        self._ws = None

    # -- reference surface ------------------------------------------------------------------
    @property
    def uses_grad_info(self) -> bool:
        raise NotImplementedError

    def get_logging_info(self):
        raise NotImplementedError

    def set_eval_mode(self, eval_setting: bool):
        raise NotImplementedError

    def create_new_point(self, x: torch.Tensor) -> Point:
        """transition_operators/base.py:30-34, on the device kernels."""
        x = x.detach().contiguous()
        log_q, gq = self._flow.cuda_log_prob(x, with_grad=self.uses_grad_info)
        n = x.shape[0]
        log_p = torch.empty(n, dtype=torch.float32, device=x.device)
        gp = torch.empty_like(x) if self.uses_grad_info else None
        rc = _lib.lib().fab_target_logprob_grad_f32(self._target.target_desc(x.device), _lib.ptr(x),
                                                    _lib.ptr(log_p), _lib.ptr(gp), n,
                                                    _lib.stream_ptr(x.device))
        _lib.check(rc, "fab_target_logprob_grad_f32")
        return Point(x, log_q, log_p, gq, gp)

    def intermediate_target_log_prob(self, point: Point, beta) -> torch.Tensor:
        g = make_gamma(beta, self.alpha, self.p_target)
        return g.cq * point.log_q + g.cp * point.log_p

    # -- helpers ------------------------------------------------------------------------------
    def _workspace(self, nbytes: int, device) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    def _world(self) -> int:
        return fdist.world(self.process_group)[0]

    def _peer_exchange(self, dev):
        """NVLink peer-memory exchange of the tuner statistics (dist.PeerExchange), or None -> NCCL.
        Created on the first multi-rank transition (collective: every rank gets here together)."""
        key = (id(self.process_group), str(dev))
        if getattr(self, "_peer_key", None) != key:
            self._peer = fdist.PeerExchange.create(self.process_group, dev)
            self._peer_key = key
        return self._peer

    @staticmethod
    def _check_point(point: Point, need_grad: bool):
        for name in ("x", "log_q", "log_p") + (("grad_log_q", "grad_log_p") if need_grad else ()):
            t = getattr(point, name)
            if t is None:
                raise ValueError(f"Point.{name} is required by this operator")
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError(f"Point.{name} must be a contiguous fp32 CUDA tensor")


class HamiltonianMonteCarlo(TransitionOperator):
    def __init__(self, n_ais_intermediate_distributions: int, dim: int,
                 base_log_prob: LogProbFunc, target_log_prob: LogProbFunc,
                 alpha: float = None, p_target: bool = False, epsilon: float = 1.0,
                 n_outer: int = 1, L: int = 5, mass_init=1.0, target_p_accept: float = 0.65,
                 max_grad: float = 1e3, tune_period: bool = False,
                 common_epsilon_init_weight: float = 0.1, eval_mode: bool = False):
        super().__init__(n_ais_intermediate_distributions, dim, base_log_prob, target_log_prob,
                         alpha=alpha, p_target=p_target)
        if isinstance(mass_init, torch.Tensor):
            assert mass_init.shape == (dim,)
        self.tune_period = tune_period
        w = common_epsilon_init_weight
        self.register_buffer("common_epsilon", torch.tensor([epsilon * w]))
        self.register_buffer("epsilons",
                             torch.ones([n_ais_intermediate_distributions, n_outer]) * epsilon * (1 - w))
        self.register_buffer("mass_vector", torch.ones(dim) * mass_init)
        self.n_outer, self.L = n_outer, L
        self.target_p_accept, self.max_grad = target_p_accept, max_grad
        self.eval_mode = eval_mode
        # device-side logging scalars: first_p_accept[n_outer], last_p_accept[n_outer],
        # avg_dist_first, avg_dist_last
        self.register_buffer("_log", torch.zeros(2 * n_outer + 4), persistent=False)
        self._seen_first = False
        self._seen_last = False
        self._prop_bufs = None

    @property
    def uses_grad_info(self) -> bool:
        return True

    def set_eval_mode(self, eval_setting: bool):
        """When eval_mode is on, no tuning of epsilon occurs (hmc.py:55-57)."""
        self.eval_mode = eval_setting

    def get_epsilon(self, i: int, n: int) -> torch.Tensor:
        return self.epsilons[i - 1, n] + self.common_epsilon        # hmc.py:90-100

    # logging views with the reference's attribute names
    @property
    def first_dist_p_accepts(self):
        return [self._log[n:n + 1].detach().cpu() for n in range(self.n_outer)]

    @property
    def last_dist_p_accepts(self):
        return [self._log[self.n_outer + n:self.n_outer + n + 1].detach().cpu()
                for n in range(self.n_outer)]

    @property
    def average_distance_first_dist(self):
        return self._log[2 * self.n_outer].detach().cpu()

    @property
    def average_distance_last_dist(self):
        if not self._seen_last:
            raise AttributeError("average_distance_last_dist")
        return self._log[2 * self.n_outer + 1].detach().cpu()

    def get_logging_info(self) -> dict:
        """Same keys as hmc.py:59-88 (one device->host copy)."""
        M, no = self.n_ais_intermediate_distributions, self.n_outer
        log = self._log.detach().cpu()
        eps = self.epsilons.detach().cpu()
        common = self.common_epsilon.detach().cpu()
        d = {}
        for n in range(no):
            d[f"dist0_p_accept_{n}"] = log[n].item()
        if M > 1:
            for n in range(no):
                d[f"dist{M - 1}_p_accept_{n}"] = log[no + n].item()
        d["epsilons_dist0_loop0"] = (eps[0 - 1, 0] + common).item()     # get_epsilon(0, 0)
        if M > 1:
            d[f"epsilons_dist{M - 1}_loop0"] = (eps[M - 2, 0] + common).item()
        d["average_distance_dist0"] = log[2 * no].item()
        if self._seen_last:
            d[f"average_distance_dist_{M - 1}"] = log[2 * no + 1].item()
        return d

    # -- launches ------------------------------------------------------------------------------
    def _state(self) -> "_lib.HmcState":
        return _lib.HmcState(_lib.ptr(self.epsilons), _lib.ptr(self.common_epsilon),
                             _lib.ptr(self.mass_vector), _lib.ptr(self._log),
                             self.n_ais_intermediate_distributions, self.n_outer)

    def _prop_point(self, like: Point, slot: int) -> Point:
        n = like.x.shape[0]
        if self._prop_bufs is None or self._prop_bufs[0].x.shape != like.x.shape or \
                self._prop_bufs[0].x.device != like.x.device:
            mk = lambda: Point(torch.empty_like(like.x), torch.empty_like(like.log_q),
                               torch.empty_like(like.log_p), torch.empty_like(like.x),
                               torch.empty_like(like.x))
            self._prop_bufs = (mk(), mk())
        return self._prop_bufs[slot]

    def chain_noise(self, M: int, n: int, d: int, dev):
        """Noise for a whole chain, one (momentum[n_outer,n,d], exponential[n_outer,n]) pair per
        transition.  With the default device RNG everything is drawn in two launches."""
        if self.chain_noise_override is not None:
            pre, self.chain_noise_override = self.chain_noise_override, None
            return pre
        if type(self.noise) is DeviceNoise:
            mom = torch.randn((M, self.n_outer, n, d), dtype=torch.float32, device=dev)
            exp = torch.empty((M, self.n_outer, n), dtype=torch.float32, device=dev).exponential_(1.0)
            return [(mom[j], exp[j]) for j in range(M)]
        return [(self.noise.momentum(j + 1, self.n_outer, n, d, dev),
                 self.noise.exponential(j + 1, self.n_outer, n, dev)) for j in range(M)]

    def chain_noise_static(self, M: int, n: int, d: int, dev):
        """Persistent noise buffers for a CUDA-graph-captured chain (ais.py: use_cuda_graph)."""
        return (torch.empty((M, self.n_outer, n, d), dtype=torch.float32, device=dev),
                torch.empty((M, self.n_outer, n), dtype=torch.float32, device=dev))

    def fill_chain_noise(self, static) -> None:
        """Refill the static buffers: injected / overridden noise is copied in, the default device
        RNG draws in place (same generator calls, in the same order, as `chain_noise`)."""
        a, b = static
        if self.chain_noise_override is not None or type(self.noise) is not DeviceNoise:
            for j, (x, y) in enumerate(self.chain_noise(a.shape[0], a.shape[2], a.shape[3], a.device)):
                a[j].copy_(x); b[j].copy_(y)
        else:
            a.normal_()
            b.exponential_(1.0)

    def run(self, point: Point, i: int, beta, log_w: Optional[torch.Tensor] = None,
            w_update=None, n_active: Optional[torch.Tensor] = None, noise=None, timings=None) -> Point:
        """One HMC transition at distribution i.  `w_update=(g_w, g_next)` (the sampler's gammas at
        beta_i and beta_{i+1}) fuses the AIS log-weight update log_w += g_next(x) - g_w(x)
        (ais.py:93-100) into the last outer step.  `noise` = pre-drawn (momentum, exponential)."""
        self._check_point(point, need_grad=True)
        flow, target = self._flow, self._target
        dev = point.x.device
        n, d = point.x.shape
        L = _lib.lib()
        g = make_gamma(beta, self.alpha, self.p_target)
        fuse_w = w_update is not None and log_w is not None
        g_w, g_next = w_update if fuse_w else (g, g)
        if noise is not None:
            mom, exp = noise
        else:
            mom = self.noise.momentum(i, self.n_outer, n, d, dev)
            exp = self.noise.exponential(i, self.n_outer, n, dev)
        tdesc = target.target_desc(dev)
        rowtile = tdesc.kind == _lib.FAB_TARGET_MANYWELL and flow.use_rowtile(n)
        if rowtile:
            ws = self._workspace(int(L.fab_umma_workspace_bytes(flow.desc(), n)), dev)
            blob = flow.umma_blob()
            step = L.fab_hmc_step_umma_f32
        else:
            ws = self._workspace(int(L.fab_hmc_workspace_bytes(flow.desc(), n)), dev)
            blob = flow.blob()
            step = L.fab_hmc_step_f32
        world = self._world()
        st = self._state()
        stream = _lib.stream_ptr(dev)
        null_pt = _lib.PointPtrs(None, None, None, None, None)
        prop_in = null_pt
        for no in range(self.n_outer):
            last = no == self.n_outer - 1
            args = _lib.HmcArgs(i, no, self.L, 0 if self.eval_mode else 1,
                                self.target_p_accept, self.max_grad, g,
                                1 if (fuse_w and last) else 0, g_w, g_next, 1 if world > 1 else 0)
            prop_out = null_pt if last else _lib.point_ptrs(self._prop_point(point, no & 1))
            if timings is not None:       # bench.py roofline: brackets the fused kernel only
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = step(flow.desc(), _lib.ptr(blob), tdesc, st, args,
                      _lib.point_ptrs(point), prop_in, prop_out,
                      _lib.ptr(log_w) if log_w is not None else None,
                      _lib.ptr(mom[no]), _lib.ptr(exp[no]),
                      _lib.ptr(n_active) if n_active is not None else None,
                      _lib.ptr(self._stats), _lib.ptr(ws), n, stream)
            _lib.check(rc, "fab_hmc_step_umma_f32" if rowtile else "fab_hmc_step_f32")
            if timings is not None:
                e1.record()
                timings.append((e0, e1))
            if world > 1:
                px = self._peer_exchange(dev)
                if px is not None:      # statistics summed over NVLink peer memory inside one small launch
                    _lib.check(L.fab_hmc_finish_peer_f32(st, args, _lib.ptr(self._stats), _lib.ptr(px.ptrs),
                                                         px.world, px.rank, _lib.ptr(px.seq), stream),
                               "fab_hmc_finish_peer_f32")
                else:
                    fdist.reduce_stats(self._stats[:4], self.process_group)
                    _lib.check(L.fab_hmc_finish_f32(st, args, _lib.ptr(self._stats), stream),
                               "fab_hmc_finish_f32")
            prop_in = prop_out
        if i == 1:
            self._seen_first = True
        elif i == self.n_ais_intermediate_distributions:
            self._seen_last = True
        return point

    def transition(self, point: Point, i: int, beta) -> Point:
        """Plugin-API entry (hmc.py:186-202): mutates and returns `point`."""
        return self.run(point, i, beta)


class Metropolis(TransitionOperator):
    def __init__(self, n_ais_intermediate_distributions: int, dim: int,
                 base_log_prob: LogProbFunc, target_log_prob: LogProbFunc, n_updates,
                 alpha: float = None, p_target: bool = False, max_step_size=1.0,
                 min_step_size=0.1, adjust_step_size=True, target_p_accept=0.65,
                 eval_mode: bool = False):
        super().__init__(n_ais_intermediate_distributions, dim, base_log_prob, target_log_prob,
                         alpha=alpha, p_target=p_target)
        if n_updates > _lib.FAB_MAX_UPDATES:
            raise ValueError(f"n_updates <= {_lib.FAB_MAX_UPDATES} supported")
        self.n_distributions = n_ais_intermediate_distributions
        self.n_updates = n_updates
        self.adjust_step_size = adjust_step_size
        self.register_buffer("noise_scalings",
                             torch.linspace(max_step_size, min_step_size, n_updates).repeat(
                                 (n_ais_intermediate_distributions, 1)))
        self.target_prob_accept = target_p_accept
        self.eval_mode = eval_mode

    @property
    def uses_grad_info(self) -> bool:
        return False

    def set_eval_mode(self, eval_setting: bool):
        self.eval_mode = not eval_setting       # inverted in the reference (quirk 4)

    def get_logging_info(self) -> Dict:
        s = self.noise_scalings.detach().cpu()
        return {"noise_scaling_0_0": s[0, 0].item(), "noise_scaling_0_-1": s[0, -1].item()}

    def chain_noise(self, M: int, n: int, d: int, dev):
        """One (proposal[n_updates,n,d], uniform[n_updates,n]) pair per transition."""
        if self.chain_noise_override is not None:
            pre, self.chain_noise_override = self.chain_noise_override, None
            return pre
        if type(self.noise) is DeviceNoise:
            prop = torch.randn((M, self.n_updates, n, d), dtype=torch.float32, device=dev)
            unif = torch.rand((M, self.n_updates, n), dtype=torch.float32, device=dev)
            return [(prop[j], unif[j]) for j in range(M)]
        return [(self.noise.proposal(j + 1, self.n_updates, n, d, dev),
                 self.noise.uniform(j + 1, self.n_updates, n, dev)) for j in range(M)]

    def chain_noise_static(self, M: int, n: int, d: int, dev):
        return (torch.empty((M, self.n_updates, n, d), dtype=torch.float32, device=dev),
                torch.empty((M, self.n_updates, n), dtype=torch.float32, device=dev))

    def fill_chain_noise(self, static) -> None:
        a, b = static
        if self.chain_noise_override is not None or type(self.noise) is not DeviceNoise:
            for j, (x, y) in enumerate(self.chain_noise(a.shape[0], a.shape[2], a.shape[3], a.device)):
                a[j].copy_(x); b[j].copy_(y)
        else:
            a.normal_()
            b.uniform_()

    def run(self, point: Point, i: int, beta, log_w: Optional[torch.Tensor] = None,
            w_update=None, n_active: Optional[torch.Tensor] = None, noise=None) -> Point:
        self._check_point(point, need_grad=False)
        flow, target = self._flow, self._target
        dev = point.x.device
        n, d = point.x.shape
        L = _lib.lib()
        g = make_gamma(beta, self.alpha, self.p_target)
        fuse_w = w_update is not None and log_w is not None
        g_w, g_next = w_update if fuse_w else (g, g)
        if noise is not None:
            prop, unif = noise
        else:
            prop = self.noise.proposal(i, self.n_updates, n, d, dev)
            unif = self.noise.uniform(i, self.n_updates, n, dev)
        ws = self._workspace(int(L.fab_metropolis_workspace_bytes(flow.desc(), n, self.n_updates)),
                             dev)
        world = self._world()
        tune = 1 if (self.adjust_step_size and not self.eval_mode) else 0
        args = _lib.MetropolisArgs(i, self.n_updates, tune, self.target_prob_accept, g,
                                   1 if fuse_w else 0, g_w, g_next, 1 if world > 1 else 0)
        pp = _lib.PointPtrs(_lib.ptr(point.x), _lib.ptr(point.log_q), _lib.ptr(point.log_p),
                            None, None)
        stream = _lib.stream_ptr(dev)
        rc = L.fab_metropolis_transition_f32(flow.desc(), _lib.ptr(flow.blob()),
                                             target.target_desc(dev), args,
                                             _lib.ptr(self.noise_scalings), pp,
                                             _lib.ptr(log_w) if log_w is not None else None,
                                             _lib.ptr(prop), _lib.ptr(unif),
                                             _lib.ptr(n_active) if n_active is not None else None,
                                             _lib.ptr(self._stats), _lib.ptr(ws), n, stream)
        _lib.check(rc, "fab_metropolis_transition_f32")
        if world > 1:
            fdist.reduce_stats(self._stats, self.process_group)
            _lib.check(L.fab_metropolis_finish_f32(args, _lib.ptr(self.noise_scalings),
                                                   _lib.ptr(self._stats), stream),
                       "fab_metropolis_finish_f32")
        return point

    def transition(self, point: Point, i: int, beta) -> Point:
        return self.run(point, i, beta)
