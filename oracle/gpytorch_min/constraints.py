"""gpytorch.constraints restated (test infrastructure): softplus-transformed
lower bounds.  Positive() == GreaterThan(0); noise uses GreaterThan(1e-4)."""
import math
import torch
from torch import nn
from torch.nn.functional import softplus


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class GreaterThan(nn.Module):
    def __init__(self, lower_bound, transform=softplus, inv_transform=inv_softplus):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(math.inf))
        self.enforced = True

    def transform(self, raw):
        return softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return inv_softplus(value - self.lower_bound)

    def check_raw(self, t):
        return True


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)
