"""Noise sources for the oracle sampler.  TEST INFRASTRUCTURE.

The reference draws its randomness straight from torch's global generators, in a
fixed order per call:

1. base sample          `torch.randn((B, d))`                  (normflows DiagGaussian.forward)
2. per HMC outer step   `torch.randn_like(x)`                  fab/.../hmc.py:134
                        `Exponential(1.).sample([B])` on CPU   fab/.../hmc.py:118
   per Metropolis step  `torch.randn(x.shape)` on CPU          fab/.../metropolis.py:57
                        `torch.rand([B])` on CPU               fab/.../metropolis.py:65

`TorchNoise` issues exactly those calls, so with the same seed the oracle consumes
the same numbers as the reference.  `RecordingNoise` additionally keeps every draw
so that the CUDA path can be fed identical noise; `ReplayNoise` plays a record back.
"""
from typing import Dict, List

import torch


class TorchNoise:
    def base_eps(self, batch: int, dim: int, dtype, device) -> torch.Tensor:
        return torch.randn((batch, dim), dtype=dtype, device=device)

    def momentum(self, x: torch.Tensor) -> torch.Tensor:
        return torch.randn_like(x)

    def exponential(self, shape, device) -> torch.Tensor:
        return torch.distributions.Exponential(1.).sample(shape).to(device)

    def proposal(self, x: torch.Tensor) -> torch.Tensor:
        return torch.randn(x.shape).to(x.device)

    def uniform(self, shape, device) -> torch.Tensor:
        return torch.rand(shape).to(device)


class RecordingNoise(TorchNoise):
    def __init__(self):
        self.record: Dict[str, List[torch.Tensor]] = {
            "base_eps": [], "momentum": [], "exponential": [], "proposal": [], "uniform": []}

    def _keep(self, key, t):
        self.record[key].append(t.detach().clone().cpu())
        return t

    def base_eps(self, batch, dim, dtype, device):
        return self._keep("base_eps", super().base_eps(batch, dim, dtype, device))

    def momentum(self, x):
        return self._keep("momentum", super().momentum(x))

    def exponential(self, shape, device):
        return self._keep("exponential", super().exponential(shape, device))

    def proposal(self, x):
        return self._keep("proposal", super().proposal(x))

    def uniform(self, shape, device):
        return self._keep("uniform", super().uniform(shape, device))


class ReplayNoise(TorchNoise):
    """Plays back a `RecordingNoise.record` (or hand-made tensors) in order."""

    def __init__(self, record: Dict[str, List[torch.Tensor]]):
        self._it = {k: iter(v) for k, v in record.items()}

    def _next(self, key, like_dtype, device):
        return next(self._it[key]).to(device=device, dtype=like_dtype)

    def base_eps(self, batch, dim, dtype, device):
        t = self._next("base_eps", dtype, device)
        assert t.shape == (batch, dim)
        return t

    def momentum(self, x):
        t = self._next("momentum", x.dtype, x.device)
        assert t.shape == x.shape, (t.shape, x.shape)
        return t

    def exponential(self, shape, device):
        return self._next("exponential", torch.get_default_dtype(), device)

    def proposal(self, x):
        return self._next("proposal", x.dtype, x.device)

    def uniform(self, shape, device):
        return self._next("uniform", torch.get_default_dtype(), device)


class Float32RecordingNoise(RecordingNoise):
    """Draws every variate in float32 (then casts to the chain dtype) so that an fp64 ground-truth
    run and the fp32 CUDA run can consume bit-identical noise."""

    def base_eps(self, batch, dim, dtype, device):
        return self._keep("base_eps", torch.randn((batch, dim), dtype=torch.float32)).to(dtype)

    def momentum(self, x):
        return self._keep("momentum", torch.randn(x.shape, dtype=torch.float32)).to(x.dtype)

    def exponential(self, shape, device):
        e = torch.empty(shape, dtype=torch.float32).exponential_(1.0)
        return self._keep("exponential", e)

    def proposal(self, x):
        return self._keep("proposal", torch.randn(x.shape, dtype=torch.float32)).to(x.dtype)

    def uniform(self, shape, device):
        return self._keep("uniform", torch.rand(shape, dtype=torch.float32))
