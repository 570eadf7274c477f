"""`Point`: struct-of-arrays chain state (fab/sampling_methods/base.py:7-47), same attributes and
mask get/set semantics as the reference so that `fab/core.py:117,167` and the trainers
(`train_with_prioritised_buffer.py:145-147`) can consume it unchanged."""
from typing import Optional

import torch


class Point:
    def __init__(self, x: torch.Tensor, log_q: torch.Tensor, log_p: torch.Tensor,
                 grad_log_q: Optional[torch.Tensor] = None,
                 grad_log_p: Optional[torch.Tensor] = None):
        self.x = x
        self.log_q = log_q
        self.log_p = log_p
        self.grad_log_q = grad_log_q
        self.grad_log_p = grad_log_p

    @property
    def device(self):
        return self.x.device

    def to(self, device):
        for name in ("x", "log_q", "log_p", "grad_log_q", "grad_log_p"):
            t = getattr(self, name)
            if t is not None:
                setattr(self, name, t.to(device))

    def __getitem__(self, indices):
        pick = lambda t: None if t is None else t[indices]
        return Point(self.x[indices], self.log_q[indices], self.log_p[indices],
                     pick(self.grad_log_q), pick(self.grad_log_p))

    def __setitem__(self, indices, values):
        self.x[indices] = values.x
        self.log_q[indices] = values.log_q
        self.log_p[indices] = values.log_p
        if self.grad_log_q is not None:
            self.grad_log_q[indices] = values.grad_log_q
            self.grad_log_p[indices] = values.grad_log_p
