"""Parameter constraints as state-carrying sub-modules (gpytorch.constraints: Positive, GreaterThan, Interval).

The transforms themselves (softplus, softplus + lower bound) are applied where the parameters are read
(kernels.py, likelihoods.py); these modules exist so that `state_dict()` carries the same
`raw_<name>_constraint.lower_bound / upper_bound` buffers as the reference's checkpoints (`state.pt`,
experiments/swiss_roll/models/exact_cmp.py:121-124) and so that `module.raw_<name>_constraint.transform(raw)`
is available to code written against gpytorch."""
import math

import torch
from torch.nn.functional import softplus

from .module import Module


class Interval(Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(float(upper_bound)))

    def transform(self, raw):
        lo, hi = self.lower_bound.to(raw.dtype), self.upper_bound.to(raw.dtype)
        return lo + (hi - lo) * torch.sigmoid(raw)

    def inverse_transform(self, value):
        lo, hi = self.lower_bound.to(value.dtype), self.upper_bound.to(value.dtype)
        p = (value - lo) / (hi - lo)
        return torch.log(p) - torch.log1p(-p)


class GreaterThan(Interval):
    def __init__(self, lower_bound):
        super().__init__(lower_bound, math.inf)

    def transform(self, raw):
        return softplus(raw) + self.lower_bound.to(raw.dtype)

    def inverse_transform(self, value):
        v = value - self.lower_bound.to(value.dtype)
        return v + torch.log(-torch.expm1(-v))


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
