"""Oracle restatement of the reference's prioritised replay buffer.  TEST INFRASTRUCTURE.

  fab/utils/prioritised_replay_buffer.py:10-17   sample_without_replacement (Gumbel-top-k)
  fab/utils/prioritised_replay_buffer.py:71-85   add (ring write)
  fab/utils/prioritised_replay_buffer.py:88-100  sample
  fab/utils/prioritised_replay_buffer.py:117-131 adjust (invalid entries kill their sample)

  fab/utils/replay_buffer.py:11-100              ReplayBuffer (rank-weighted multinomial without
                                                 replacement over the ring; OracleReplayBuffer below)

Pinned against the unmodified reference class by oracle/gen_golden.py (same seeds -> same
buffer contents, same sampled index set, same state after adjust); the Gumbel noise can be
injected so that the CUDA path is compared on identical numbers.
"""
from typing import Optional, Tuple

import torch


def gumbel_like(logits: torch.Tensor) -> torch.Tensor:
    """The reference's draw (:12-13): a CPU Gumbel(0,1) sample of logits.shape."""
    return torch.distributions.Gumbel(torch.tensor(0.0), torch.tensor(1.0)).sample(logits.shape)


def topk_set(logits: torch.Tensor, z: torch.Tensor, n: int) -> torch.Tensor:
    """Sorted indices of the n largest z + logits (the reference's topk is unsorted and then
    permuted; as a set it is this)."""
    return torch.sort(torch.topk(z + logits, n, sorted=False).indices).values


class OracleBuffer:
    def __init__(self, dim: int, max_length: int, min_sample_length: int):
        assert min_sample_length < max_length
        self.dim, self.max_length, self.min_sample_length = dim, max_length, min_sample_length
        self.x = torch.zeros(max_length, dim)
        self.log_w = torch.zeros(max_length)
        self.log_q_old = torch.zeros(max_length)
        self.current_index, self.is_full, self.can_sample = 0, False, False

    def add(self, x, log_w, log_q_old) -> None:
        b = x.shape[0]
        idx = (torch.arange(b) + self.current_index) % self.max_length
        self.x[idx] = x
        self.log_w[idx] = log_w
        self.log_q_old[idx] = log_q_old
        new_index = self.current_index + b
        if not self.is_full:
            self.is_full = new_index >= self.max_length
            self.can_sample = new_index >= self.min_sample_length
        self.current_index = new_index % self.max_length

    def sample(self, batch_size: int, z: Optional[torch.Tensor] = None
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        if not self.can_sample:
            raise Exception("Buffer must be at minimum length before calling sample")
        max_index = self.max_length if self.is_full else self.current_index
        logits = self.log_w[:max_index]
        if z is None:
            z = gumbel_like(logits)
        idx = topk_set(logits, z, batch_size)
        return self.x[idx], self.log_w[idx], self.log_q_old[idx], idx

    def adjust(self, log_w_adjustment, log_q, indices) -> None:
        valid = torch.isfinite(log_w_adjustment) & torch.isfinite(log_q)
        self.log_w[indices[valid]] += log_w_adjustment[valid]
        self.log_q_old[indices[valid]] = log_q[valid]
        self.log_w[indices[~valid]] = -float("inf")


# ---- fab/utils/replay_buffer.py: the un-prioritised buffer -------------------------------------
def exponential_like(probs: torch.Tensor) -> torch.Tensor:
    """The draw inside `torch.multinomial(probs, k, replacement=False)` on the CPU (:84): one
    Exponential(1) variate per entry from the global generator (ATen: q = empty_like(probs)
    .exponential_(1); topk(probs / q) -- checked against torch.multinomial by the golden script)."""
    return torch.empty_like(probs).exponential_(1)


def race_topk(probs: torch.Tensor, q: torch.Tensor, n: int) -> torch.Tensor:
    """Indices of the n largest probs / q in descending order = multinomial without replacement."""
    return torch.topk(probs / q, n).indices


class OracleReplayBuffer:
    """fab/utils/replay_buffer.py:11-100 without the device plumbing; `sample` takes the exponential
    variates so that the CUDA path can be compared on identical numbers."""

    def __init__(self, dim: int, max_length: int, min_sample_length: int, temperature: float = 1.0):
        assert min_sample_length < max_length
        self.dim, self.max_length, self.min_sample_length = dim, max_length, min_sample_length
        self.x = torch.zeros(max_length, dim)
        self.log_w = torch.zeros(max_length)
        self.add_count = torch.zeros(max_length)
        self.current_index, self.current_add_count = 0, 0
        self.is_full, self.can_sample = False, False
        self.temperature = temperature

    def fill(self, initial_sampler) -> None:                 # :52-57
        while self.can_sample is False:
            x, log_w = initial_sampler()
            self.add(x, log_w)
            self.current_add_count = 0
        self.current_add_count = 1

    def add(self, x, log_w) -> None:                          # :59-74
        b = x.shape[0]
        idx = (torch.arange(b) + self.current_index) % self.max_length
        self.x[idx] = x
        self.log_w[idx] = log_w
        self.add_count[idx] = self.current_add_count
        new_index = self.current_index + b
        if not self.is_full:
            self.is_full = new_index >= self.max_length
            self.can_sample = new_index >= self.min_sample_length
        self.current_index = new_index % self.max_length
        self.current_add_count += 1

    def probs(self) -> torch.Tensor:                          # :81-83
        max_index = self.max_length if self.is_full else self.current_index
        rank = self.current_add_count - self.add_count[:max_index]
        return torch.pow(1 / rank, self.temperature)

    def sample(self, batch_size: int, q: Optional[torch.Tensor] = None):   # :76-86
        if not self.can_sample:
            raise Exception("Buffer must be at minimum length before calling sample")
        probs = self.probs()
        if q is None:
            q = exponential_like(probs)
        idx = race_topk(probs, q, batch_size)
        return self.x[idx], self.log_w[idx], idx
