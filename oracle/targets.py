"""Oracle restatement of the target densities on the hot path.  TEST INFRASTRUCTURE.

  fab/target_distributions/many_well.py:81-90 + double_well.py:44-58  ManyWellEnergy.log_prob
  fab/target_distributions/many_well.py:53-54, double_well.py:97-103   log Z
  fab/target_distributions/gmm.py:13-66                                GMM ctor / log_prob
Pinned against the reference by oracle/gen_golden.py and against the published
constants (log Z = 164.69567532 at d=32; first GMM-40 means) in tests/test_oracle.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

DOUBLE_WELL_Z = 11784.50927   # double_well.py:68,99


class OracleManyWell:
    """d/2 independent copies of the 2-D double well: E(x1,x2) = a x1 + b x1^2 + c x1^4 + x2^2/2."""

    def __init__(self, dim: int = 4, a: float = -0.5, b: float = -6.0, c: float = 1.0,
                 normalised: bool = False):
        assert dim % 2 == 0
        self.dim, self.n_wells = dim, dim // 2
        self.a, self.b, self.c = a, b, c
        self.normalised = normalised

    @property
    def log_Z(self) -> torch.Tensor:
        assert (self.a, self.b, self.c) == (-0.5, -6.0, 1.0)
        return torch.tensor((np.log(DOUBLE_WELL_Z) + 0.5 * np.log(2 * torch.pi)) * self.n_wells)

    def _pair_log_prob(self, xy: torch.Tensor) -> torch.Tensor:
        x1, x2 = xy[:, 0], xy[:, 1]
        e1 = self.a * x1 + self.b * x1.pow(2) + self.c * x1.pow(4)
        e2 = 0.5 * x2.pow(2)
        return torch.squeeze(-((e1 + e2) / 1.))

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        lp = torch.sum(torch.stack([self._pair_log_prob(x[:, 2 * i:2 * i + 2])
                                    for i in range(self.n_wells)]), dim=0)
        return lp - self.log_Z if self.normalised else lp

    def grad_log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """Closed-form gradient (what the CUDA kernel computes; the reference uses autograd)."""
        g = torch.empty_like(x)
        x1 = x[:, 0::2]
        g[:, 0::2] = -(self.a + 2 * self.b * x1 + 4 * self.c * x1 ** 3)
        g[:, 1::2] = -x[:, 1::2]
        return g


class OracleGMM:
    """Equal-weight mixture of isotropic Gaussians; parameters are a pure function of the
    torch global RNG state at construction (gmm.py:22: `torch.rand((n_mixes, dim))`)."""

    def __init__(self, dim: int, n_mixes: int, loc_scaling: float, log_var_scaling: float = 0.1):
        self.dim, self.n_mixes = dim, n_mixes
        self.locs = (torch.rand((n_mixes, dim)) - 0.5) * 2 * loc_scaling
        log_var = torch.ones((n_mixes, dim)) * log_var_scaling
        self.scale_trils = torch.diag_embed(F.softplus(log_var))
        self.cat_probs = torch.ones(n_mixes)

    @property
    def distribution(self):
        mix = torch.distributions.Categorical(self.cat_probs)
        com = torch.distributions.MultivariateNormal(self.locs, scale_tril=self.scale_trils,
                                                     validate_args=False)
        return torch.distributions.MixtureSameFamily(mixture_distribution=mix,
                                                     component_distribution=com,
                                                     validate_args=False)

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        lp = self.distribution.log_prob(x)
        mask = torch.zeros_like(lp)
        mask[lp < -1e4] = -torch.tensor(float("inf"))
        return lp + mask

    def sample(self, shape=(1,)):
        return self.distribution.sample(shape)


class OracleDiagGaussian:
    """Stand-in for the `WrappedTorchDist(MultivariateNormal(loc, s*I))` base/target the
    reference tests use (fab/sampling_methods/ais_test.py:30-33,96-97; fab/wrappers/torch.py)."""

    def __init__(self, loc: torch.Tensor, scale: float):
        self._d = torch.distributions.MultivariateNormal(
            loc=loc, scale_tril=scale * torch.eye(loc.shape[0], dtype=loc.dtype))
        self.dim = loc.shape[0]

    def sample_and_log_prob(self, shape):
        x = self._d.sample(shape)
        return x, self._d.log_prob(x)

    def sample(self, shape):
        return self._d.sample(shape)

    def log_prob(self, x):
        return self._d.log_prob(x)

    @property
    def event_shape(self):
        return (self.dim,)


def aldp_surrogate_tables(dim: int = 60, seed: int = 0):
    """Parameters of the ALDP surrogate (BASELINE config 5; build-defined, the reference's
    AldpBoltzmann -- fab/target_distributions/aldp.py:17-159 -- needs OpenMM and cannot run
    here): every third coordinate is a torsion with multiplicity 1..3, the others are harmonic
    (bond/angle-like, in the normalised internal coordinates the reference's flow works in); the
    first two torsions are coupled like phi/psi.  A pure function of (dim, seed)."""
    g = torch.Generator().manual_seed(1000 + seed)
    tors = torch.arange(dim) % 3 == 2
    mult = torch.where(tors, torch.randint(1, 4, (dim,), generator=g).float(), torch.zeros(dim))
    p0 = torch.where(tors, 0.5 + 2.5 * torch.rand(dim, generator=g), 1.0 + 24.0 * torch.rand(dim, generator=g))
    p1 = torch.where(tors, (torch.rand(dim, generator=g) * 2 - 1) * math.pi, torch.rand(dim, generator=g) - 0.5)
    idx = torch.nonzero(tors).flatten()
    return dict(mult=mult, p0=p0, p1=p1, ia=int(idx[0]), ib=int(idx[1]), coupling=1.5)


class OracleAldpSurrogate:
    """E(x) = sum_harmonic k/2 (x-m)^2 + sum_torsion A (1 - cos(n x - phi)) + C (1 - cos(x_ia - x_ib));
    log_prob = -E.  Plain torch ops (autograd gives the gradient the kernel has in closed form)."""

    def __init__(self, dim: int = 60, seed: int = 0, dtype=torch.float32):
        t = aldp_surrogate_tables(dim, seed)
        self.dim = dim
        self.mult, self.p0, self.p1 = (t[k].to(dtype) for k in ("mult", "p0", "p1"))
        self.ia, self.ib, self.coupling = t["ia"], t["ib"], t["coupling"]

    def double(self):
        self.mult, self.p0, self.p1 = self.mult.double(), self.p0.double(), self.p1.double()
        return self

    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        harm = 0.5 * self.p0 * (x - self.p1) ** 2
        tors = self.p0 * (1 - torch.cos(self.mult * x - self.p1))
        e = torch.where(self.mult == 0, harm, tors).sum(dim=1)
        e = e + self.coupling * (1 - torch.cos(x[:, self.ia] - x[:, self.ib]))
        return -e


def to_double(target):
    """fp64 copy of an OracleGMM's tables (ground-truth runs)."""
    if isinstance(target, OracleGMM):
        target.locs = target.locs.double()
        target.scale_trils = target.scale_trils.double()
        target.cat_probs = target.cat_probs.double()
    return target
