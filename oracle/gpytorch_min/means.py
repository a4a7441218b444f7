"""gpytorch.means restated (test infrastructure)."""
import torch
from .module import Module


class Mean(Module):
    def __call__(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        return self.forward(x)


class ZeroMean(Mean):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)


class ConstantMean(Mean):
    def __init__(self):
        super().__init__()
        self.register_parameter("constant", torch.nn.Parameter(torch.zeros(1)))

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])
