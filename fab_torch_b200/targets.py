"""Target densities of the hot path with closed-form gradients evaluated on the device.

  ManyWellEnergy  fab/target_distributions/many_well.py:16-90 (+ double_well.py:44-58,97-103)
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


class ManyWellEnergy(_DeviceTarget):
    """d/2 copies of the 2-D double well E = a x1 + b x1^2 + c x1^4 + x2^2/2."""

    def __init__(self, dim: int = 4, use_gpu: bool = True, normalised: bool = False,
                 a: float = -0.5, b: float = -6.0, c: float = 1.0):
        super().__init__()
        assert dim % 2 == 0
        self.dim = dim
        self.n_wells = dim // 2
        self._a, self._b, self._c = a, b, c
        self.normalised = normalised
        self.device = "cuda" if (use_gpu and torch.cuda.is_available()) else "cpu"

    @property
    def log_Z_2D(self):
        if self._a == -0.5 and self._b == -6 and self._c == 1.0:
            return np.log(11784.50927) + 0.5 * np.log(2 * torch.pi)   # double_well.py:97-103
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


class GMM(_DeviceTarget):
    """Equal-weight mixture of `n_mixes` isotropic Gaussians.  Construction draws the means with
    `torch.rand((n_mixes, dim))` exactly like gmm.py:22, so the same seed gives the same target."""

    def __init__(self, dim, n_mixes, loc_scaling, log_var_scaling=0.1, seed=0, use_gpu=True):
        super().__init__()
        self.seed, self.n_mixes, self.dim = seed, n_mixes, dim
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
