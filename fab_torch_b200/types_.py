"""Plugin-boundary ABCs, mirroring the reference so that objects built here are accepted wherever
the reference expects its own (duck typing; `fab/core.py` never isinstance-checks):

  Distribution            fab/types_.py:8-27
  TrainableDistribution   fab/trainable_distributions/base.py:4
  TargetDistribution      fab/target_distributions/base.py:7-36
"""
import abc
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn

LogProbFunc = Callable[[torch.Tensor], torch.Tensor]


class Distribution(abc.ABC):
    @abc.abstractmethod
    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @abc.abstractmethod
    def sample_and_log_prob(self, shape: Tuple) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    @abc.abstractmethod
    def sample(self, shape: Tuple) -> torch.Tensor:
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def event_shape(self) -> Tuple[int, ...]:
        raise NotImplementedError


class TrainableDistribution(Distribution, nn.Module):
    """Base class for trainable distributions."""


class TargetDistribution(abc.ABC):
    @abc.abstractmethod
    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        """(unnormalised) log probability of samples x"""
        raise NotImplementedError

    def performance_metrics(self, samples: torch.Tensor, log_w: torch.Tensor,
                            log_q_fn: Optional[LogProbFunc] = None,
                            batch_size: Optional[int] = None) -> Dict:
        raise NotImplementedError

    def sample(self, shape):
        raise NotImplementedError
