"""Target densities of the hot path with closed-form gradients evaluated on the device.

  ManyWellEnergy  fab/target_distributions/many_well.py:16-147 (+ double_well.py:44-103)
  GMM             fab/target_distributions/gmm.py:12-66
  DiagGaussianTarget  stand-in for the `WrappedTorchDist(MultivariateNormal(loc, s*I))` targets of
                  the reference's own tests (fab/sampling_methods/ais_test.py:32-33)

Each exposes `log_prob(x)` (differentiable w.r.t. x through the kernel's analytic gradient, so the
reference's `grad_and_value`, fab/sampling_methods/base.py:50-56, works on it) and
`target_desc()` -> the C-ABI descriptor the fused kernels consume.
"""
import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from fab_torch_b200 import _lib
from fab_torch_b200.types_ import TargetDistribution


class _TargetLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, x):
        x_ = _lib.f32(x.detach()).contiguous()
        n = x_.shape[0]
        log_p = torch.empty(n, dtype=torch.float32, device=x_.device)
        grad = torch.empty_like(x_) if x.requires_grad else None
        desc = target.target_desc(x_.device)
        rc = _lib.lib().fab_target_logprob_grad_f32(desc, _lib.ptr(x_), _lib.ptr(log_p),
                                                    _lib.ptr(grad), n, _lib.stream_ptr(x_.device))
        _lib.check(rc, "fab_target_logprob_grad_f32")
        ctx.save_for_backward(grad if grad is not None else x_.new_empty(0))
        return log_p

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return None, (g[:, None] * grad if grad.numel() else None)


class _DeviceTarget(nn.Module, TargetDistribution):
    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        return _TargetLogProb.apply(self, x)

    def target_desc(self, device) -> "_lib.TargetDesc":
        raise NotImplementedError


class _ChunkedDataset:
    """Batches of a test set (fab/utils/training.py:36-53).  Like the reference it yields chunks of
    size ceil(N / batch_size) -- `torch.split` is given the number of splits as the chunk size."""

    def __init__(self, batch_size: int, dataset: torch.Tensor, device):
        self.batch_size = batch_size
        self.n_splits = int(np.ceil(dataset.shape[0] / batch_size))
        self._chunks = iter(torch.split(dataset, self.n_splits))
        self.device = device
        self.test_set_n_points = dataset.shape[0]

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._chunks).to(self.device)

    def __len__(self):
        return self.n_splits


class ManyWellEnergy(_DeviceTarget):
    """d/2 copies of the 2-D double well E = a x1 + b x1^2 + c x1^4 + x2^2/2.

    Hot path: `log_prob` / `target_desc` (CUDA kernels).  Evaluation path (host
    logic in torch on `self.device`, same RNG calls in the same order as the reference):
    `sample` (many_well.py:61-67, double_well.py:60-94, rejection_sampling.py:6-20),
    `get_modes_test_set_iterator` (many_well.py:69-79) and `performance_metrics`
    (many_well.py:96-147), which is what `FABModel.get_eval_info` (fab/core.py:191-220) calls."""

    centre = 1.7
    max_dim_for_all_modes = 40          # a full mode test set has 2^(d/2) rows
    _Z_FIRST_DIM = 11784.50927          # double_well.py:68

    def __init__(self, dim: int = 4, use_gpu: bool = True, normalised: bool = False,
                 a: float = -0.5, b: float = -6.0, c: float = 1.0):
        super().__init__()
        assert dim % 2 == 0
        self.dim = dim
        self.n_wells = dim // 2
        self._a, self._b, self._c = a, b, c
        self.normalised = normalised
        if self._default_well:          # proposal of the exact sampler (double_well.py:38-41)
            self.register_buffer("component_mix", torch.tensor([0.2, 0.8]))
            self.register_buffer("means", torch.tensor([-1.7, 1.7]))
            self.register_buffer("scales", torch.tensor([0.5, 0.5]))
        if dim < self.max_dim_for_all_modes:
            # every sign pattern of +-centre on the first coordinate of each well, first well
            # slowest (the reference's meshgrid order, many_well.py:27-35)
            n = self.n_wells
            bits = (torch.arange(2 ** n)[:, None] >> torch.arange(n - 1, -1, -1)[None, :]) & 1
            modes = torch.zeros((2 ** n, dim))
            modes[:, 0::2] = (2.0 * bits - 1.0) * self.centre
            self.register_buffer("_test_set_modes", modes)
        self.shallow_well_bounds = [-1.75, -1.65]
        self.deep_well_bounds = [1.7, 1.8]
        self.device = "cuda" if (use_gpu and torch.cuda.is_available()) else "cpu"
        if self.device == "cuda":
            self.cuda()

    @property
    def _default_well(self) -> bool:
        return self._a == -0.5 and self._b == -6 and self._c == 1.0

    @property
    def log_Z_2D(self):
        if self._default_well:
            return np.log(self._Z_FIRST_DIM) + 0.5 * np.log(2 * torch.pi)   # double_well.py:97-103
        raise NotImplementedError

    @property
    def log_Z(self):
        return torch.tensor(self.log_Z_2D * self.n_wells)

    @property
    def Z(self):
        return torch.exp(self.log_Z)

    def target_desc(self, device=None):
        return _lib.TargetDesc(_lib.FAB_TARGET_MANYWELL, self.dim, 0, 0, self._a, self._b, self._c,
                               float(self.log_Z) if self.normalised else 0.0, None, None, None)

    # ------------------------------------------------------------------ evaluation path
    def _sample_first_dimension(self, n: int) -> torch.Tensor:
        """Rejection sampling of exp(-x^4 + 6x^2 + x/2) under 3 Z * (0.2 N(-1.7,.5) + 0.8 N(1.7,.5)):
        draw 10x the missing count, keep the accepted ones in order, repeat for the remainder."""
        if not self._default_well:
            raise NotImplementedError
        k = self._Z_FIRST_DIM * 3
        prop = torch.distributions.MixtureSameFamily(
            mixture_distribution=torch.distributions.Categorical(self.component_mix),
            component_distribution=torch.distributions.Normal(self.means, self.scales))
        kept, missing = [], n
        while missing > 0:
            z = prop.sample((missing * 10,))
            u = torch.distributions.Uniform(0, k * torch.exp(prop.log_prob(z))).sample().to(z)
            log_t = -(z ** 4)
            log_t = log_t + 6 * z ** 2
            log_t = log_t + 1 / 2 * z
            ok = z[torch.exp(log_t) > u][:missing]
            kept.append(ok)
            missing -= ok.shape[0]
        return kept[0] if len(kept) == 1 else torch.concat(kept, dim=0)

    def sample(self, shape):
        """Exact samples: per well, x1 by rejection sampling, x2 ~ N(0, 1)."""
        assert len(shape) == 1
        wells = []
        for _ in range(self.n_wells):
            x1 = self._sample_first_dimension(shape[0])
            x2 = torch.distributions.Normal(torch.tensor(0.0).to(x1.device),
                                            torch.tensor(1.0).to(x1.device)).sample(shape)
            wells.append(torch.stack([x1, x2], dim=-1))
        return torch.concat(wells, dim=-1)

    def get_modes_test_set_iterator(self, batch_size: int):
        """Points placed at the modes: all of them below 40 dimensions, 10^4 random ones above."""
        if self.dim < self.max_dim_for_all_modes:
            test_set = self._test_set_modes
        else:
            outer = int(1e4)
            test_set = torch.zeros((outer, self.dim))
            test_set[:, torch.arange(self.dim) % 2 == 0] = \
                -self.centre + self.centre * 2 * torch.randint(high=2, size=(outer, int(self.dim / 2)))
        return _ChunkedDataset(batch_size=batch_size, dataset=test_set, device=self.device)

    def performance_metrics(self, samples: torch.Tensor, log_w: torch.Tensor,
                            log_q_fn=None, batch_size: Optional[int] = None):
        """log-Z accuracy of 50 interleaved sub-estimates and, given the model density, its mean
        log-likelihood on the mode test set and on exact samples + the forward KL.  Quirks of the
        reference kept: `split(50)` chunks have SIZE 50; the number of exact-sample batches is
        read off the stacked tensor's first axis, i.e. max(50 // batch_size, 1)."""
        del samples
        n_runs = 50
        keep = (log_w.shape[0] // n_runs) * n_runs
        stacked = torch.stack(log_w[:keep].split(n_runs), dim=-1)
        log_Z_estimate = torch.logsumexp(stacked, dim=-1) - np.log(stacked.shape[-1])
        relative_error = torch.exp(log_Z_estimate - self.log_Z) - 1
        info = dict(relative_MSE_Z_estimate=torch.mean(torch.abs(relative_error)).cpu().item(),
                    abs_MSE_log_Z_estimate=torch.mean(torch.abs(log_Z_estimate - self.log_Z)).cpu().item())
        if log_q_fn is None:
            return info
        assert batch_size is not None
        n_batches = max(stacked.shape[0] // batch_size, 1)
        sum_modes, sum_exact, sum_kl = 0.0, 0.0, 0.0
        modes = self.get_modes_test_set_iterator(batch_size=batch_size)
        for x in modes:
            sum_modes += torch.sum(log_q_fn(x)).detach().cpu()
        for _ in range(n_batches):
            x = self.sample((batch_size,))
            log_q = log_q_fn(x)
            sum_exact += torch.sum(log_q).detach().cpu()
            sum_kl += torch.sum(self.log_prob(x) - self.log_Z - log_q).detach().cpu()
        m = batch_size * n_batches
        info.update(test_set_modes_mean_log_prob=(sum_modes / modes.test_set_n_points).cpu().item(),
                    test_set_exact_mean_log_prob=(sum_exact / m).cpu().item(),
                    forward_kl=(sum_kl / m).cpu().item(), eval_batch_size=m)
        return info


class GMM(_DeviceTarget):
    """Equal-weight mixture of `n_mixes` isotropic Gaussians.  Construction draws the means with
    `torch.rand((n_mixes, dim))` exactly like gmm.py:22, so the same seed gives the same target."""

    # evaluation path (class-level defaults so that subclasses with their own constructor have them)
    n_test_set_samples = 1000
    true_expectation_estimation_n_samples = int(1e7)
    _true_expectation = None
    _quad_tables = None

    def __init__(self, dim, n_mixes, loc_scaling, log_var_scaling=0.1, seed=0,
                 n_test_set_samples=1000, use_gpu=True,
                 true_expectation_estimation_n_samples=int(1e7)):
        """Same signature as gmm.py:13-15.  Unlike the reference, the Monte-Carlo estimate of the
        test function's true expectation (10^7 samples by default, and a re-seeding of the global
        RNG as a side effect) is NOT run here but on first use (`true_expectation`)."""
        super().__init__()
        self.seed, self.n_mixes, self.dim = seed, n_mixes, dim
        self.n_test_set_samples = n_test_set_samples
        self.true_expectation_estimation_n_samples = true_expectation_estimation_n_samples
        mean = (torch.rand((n_mixes, dim)) - 0.5) * 2 * loc_scaling
        log_var = torch.ones((n_mixes, dim)) * log_var_scaling
        self.register_buffer("cat_probs", torch.ones(n_mixes))
        self.register_buffer("locs", mean)
        self.register_buffer("scale_trils", torch.diag_embed(F.softplus(log_var)))
        self.mask_below_1e4 = True
        if use_gpu and torch.cuda.is_available():
            self.cuda()
        self._cache = None

    def _tables(self, device):
        key = (self.locs.data_ptr(), self.locs._version, str(device))
        if self._cache is None or self._cache[0] != key:
            locs = self.locs.to(device=device, dtype=torch.float32).contiguous()
            scales = torch.diagonal(self.scale_trils, dim1=-2, dim2=-1).to(
                device=device, dtype=torch.float32).contiguous()
            logw = torch.log_softmax(torch.log(self.cat_probs.to(device=device,
                                                                 dtype=torch.float32)), dim=0).contiguous()
            self._cache = (key, locs, scales, logw)
        return self._cache[1:]

    def target_desc(self, device=None):
        device = device if device is not None else self.locs.device
        locs, scales, logw = self._tables(device)
        return _lib.TargetDesc(_lib.FAB_TARGET_GMM, self.dim, self.n_mixes,
                               1 if self.mask_below_1e4 else 0, 0.0, 0.0, 0.0, 0.0,
                               _lib.ptr(locs), _lib.ptr(scales), _lib.ptr(logw))

    @property
    def distribution(self):
        mix = torch.distributions.Categorical(self.cat_probs)
        com = torch.distributions.MultivariateNormal(self.locs, scale_tril=self.scale_trils,
                                                     validate_args=False)
        return torch.distributions.MixtureSameFamily(mixture_distribution=mix,
                                                     component_distribution=com,
                                                     validate_args=False)

    def sample(self, shape=(1,)):
        return self.distribution.sample(shape)

    # ------------------------------------------------------------------ evaluation path
    # gmm.py:53-55,71-99 + fab/utils/numerical.py:8-15,25-60 (host logic in torch; the densities
    # themselves go through the CUDA kernels)
    @property
    def test_set(self) -> torch.Tensor:
        """A FRESH sample set on every access, like the reference's property."""
        return self.sample((self.n_test_set_samples,))

    def expectation_function(self, x: torch.Tensor) -> torch.Tensor:
        """The reference's quadratic test function (x+s)^T A (x+s) + b^T (x+s) with s = 2 randn(d),
        A = 2 rand(d,d), b = rand(d) drawn after seeding with 0 -- here from a local generator, so
        the values are the reference's but the global RNG is left alone."""
        if self._quad_tables is None or self._quad_tables[0].shape[0] != x.shape[-1]:
            g = torch.Generator().manual_seed(0)
            d = x.shape[-1]
            self._quad_tables = (2 * torch.randn(d, generator=g), 2 * torch.rand((d, d), generator=g),
                                 torch.rand(d, generator=g))
        s, A, b = (t.to(x) for t in self._quad_tables)
        x = x + s
        return torch.einsum("bi,ij,bj->b", x, A, x) + torch.einsum("i,bi->b", b, x)

    @property
    def true_expectation(self) -> torch.Tensor:
        if self._true_expectation is None:
            x = self.distribution.sample((self.true_expectation_estimation_n_samples,))
            self._true_expectation = torch.mean(self.expectation_function(x))
        return self._true_expectation

    def evaluate_expectation(self, samples: torch.Tensor, log_w: torch.Tensor) -> torch.Tensor:
        weights = torch.softmax(log_w, dim=-1)
        estimate = weights @ self.expectation_function(samples)
        truth = self.true_expectation.to(estimate.device)
        return (estimate - truth) / truth

    def performance_metrics(self, samples: torch.Tensor, log_w: torch.Tensor, log_q_fn=None,
                            batch_size: Optional[int] = None):
        bias_normed = self.evaluate_expectation(samples, log_w)
        bias_no_correction = self.evaluate_expectation(samples, torch.ones_like(log_w))
        if not log_q_fn:
            return {"bias_normed": bias_normed.cpu().item(),
                    "bias_no_correction": torch.abs(bias_no_correction).cpu().item()}
        log_q_test = log_q_fn(self.test_set)            # (two different test sets: reference quirk)
        log_p_test = self.log_prob(self.test_set)
        ratio = log_p_test - log_q_test
        return {"test_set_mean_log_prob": torch.mean(log_q_test).cpu().item(),
                "bias_normed": torch.abs(bias_normed).cpu().item(),
                "bias_no_correction": torch.abs(bias_no_correction).cpu().item(),
                "ess_over_p": (1 / torch.mean(torch.exp(ratio))).detach().cpu().item(),
                "kl_forward": torch.mean(ratio).detach().cpu().item()}


class DiagGaussianTarget(GMM):
    """Single diagonal Gaussian N(loc, diag(scale)^2) as a one-component 'mixture' without the
    -1e4 mask."""

    def __init__(self, loc: torch.Tensor, scale, use_gpu=True):
        nn.Module.__init__(self)
        dim = loc.shape[0]
        self.seed, self.n_mixes, self.dim = 0, 1, dim
        sc = torch.as_tensor(scale, dtype=torch.float32) * torch.ones(dim)
        self.register_buffer("cat_probs", torch.ones(1))
        self.register_buffer("locs", loc.reshape(1, dim).to(torch.float32))
        self.register_buffer("scale_trils", torch.diag_embed(sc.reshape(1, dim)))
        self.mask_below_1e4 = False
        if use_gpu and torch.cuda.is_available():
            self.cuda()
        self._cache = None


class AldpSurrogateEnergy(_DeviceTarget):
    """Closed-form 60-dof stand-in for the reference's OpenMM alanine-dipeptide density
    (fab/target_distributions/aldp.py:17-159, BASELINE config 5): harmonic bond/angle-like
    coordinates, periodic torsions with multiplicity 1..3 and one phi/psi-style coupling
    (include/fab_b200.h: FAB_TARGET_ALDP_SURROGATE).  The tables are a pure function of
    (dim, seed); `tables` may be passed to share them with another implementation."""

    def __init__(self, dim: int = 60, seed: int = 0, tables=None, use_gpu: bool = True):
        super().__init__()
        if tables is None:
            g = torch.Generator().manual_seed(1000 + seed)
            tors = torch.arange(dim) % 3 == 2
            mult = torch.where(tors, torch.randint(1, 4, (dim,), generator=g).float(), torch.zeros(dim))
            p0 = torch.where(tors, 0.5 + 2.5 * torch.rand(dim, generator=g),
                             1.0 + 24.0 * torch.rand(dim, generator=g))
            p1 = torch.where(tors, (torch.rand(dim, generator=g) * 2 - 1) * math.pi,
                             torch.rand(dim, generator=g) - 0.5)
            idx = torch.nonzero(tors).flatten()
            tables = dict(mult=mult, p0=p0, p1=p1, ia=int(idx[0]), ib=int(idx[1]), coupling=1.5)
        self.dim = dim
        self.register_buffer("mult", tables["mult"].float().contiguous())
        self.register_buffer("p0", tables["p0"].float().contiguous())
        self.register_buffer("p1", tables["p1"].float().contiguous())
        self.ia, self.ib, self.coupling = int(tables["ia"]), int(tables["ib"]), float(tables["coupling"])
        if use_gpu and torch.cuda.is_available():
            self.cuda()

    def target_desc(self, device=None):
        if device is not None and self.p0.device != torch.device(device):
            self.to(device)
        return _lib.TargetDesc(_lib.FAB_TARGET_ALDP_SURROGATE, self.dim, 0, 0, self.coupling,
                               float(self.ia), float(self.ib), 0.0, _lib.ptr(self.p1),
                               _lib.ptr(self.p0), _lib.ptr(self.mult))
