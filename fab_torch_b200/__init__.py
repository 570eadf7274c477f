"""fab_torch_b200: B200-native (sm_100a) implementation of fab-torch's annealed-importance-sampling
hot path behind the reference's own plugin API.  See DESIGN.md / INTEGRATION.md."""
from fab_torch_b200.point import Point
from fab_torch_b200.types_ import Distribution, TrainableDistribution, TargetDistribution
from fab_torch_b200.flow import B200RealNVP, make_wrapped_b200_realnvp
from fab_torch_b200.targets import ManyWellEnergy, GMM, DiagGaussianTarget, AldpSurrogateEnergy
from fab_torch_b200.transition_operators import (TransitionOperator, HamiltonianMonteCarlo,
                                                  Metropolis, DeviceNoise, InjectedNoise,
                                                  make_gamma)
from fab_torch_b200.ais import AnnealedImportanceSampler, LoggingInfo, setup_distribution_spacing
from fab_torch_b200.numerical import effective_sample_size
from fab_torch_b200.resample import (systematic_resample, systematic_ancestors,
                                      global_systematic_resample, resample_if_ess_below)

from fab_torch_b200.replay_buffer import PrioritisedReplayBuffer, ReplayData, ReplayBuffer, AISData

__version__ = "0.1"
