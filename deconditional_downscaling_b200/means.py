"""Mean modules: gpytorch's ZeroMean / ConstantMean surface plus the reference's
CMEAggregateMean (src/means/cme_aggregate_mean.py:4-55)."""
import torch
from torch import nn

from . import functional as F
from .kernels import _as_2d, cme_weights
from .module import Module


class Mean(Module):
    def __call__(self, x):
        return self.forward(_as_2d(x))


class ZeroMean(Mean):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)


class ConstantMean(Mean):
    def __init__(self):
        super().__init__()
        self.register_parameter("constant", nn.Parameter(torch.zeros(1)))

    def forward(self, x):
        return self.constant.to(x.dtype).expand(x.shape[:-1])


class CMEAggregateMean(Mean):
    """m = B^T Lambda^{-1} mu with B = l(bags_values, x)  (reference: (B^T R)(R^T mu),
    src/means/cme_aggregate_mean.py:35-55).  Computed as W^T mu with the W = Lambda^{-1} B that
    CMEAggregateKernel needs anyway (shared through the KernelInverse cache)."""

    def __init__(self, bag_kernel, bags_values, individuals_mean, root_inv_bags_covar):
        super().__init__()
        self.bag_kernel = bag_kernel
        self.register_buffer("bags_values", bags_values)
        self.register_buffer("individuals_mean", individuals_mean)
        self.root_inv_bags_covar = root_inv_bags_covar

    def forward(self, x):
        W = cme_weights(self.root_inv_bags_covar, self.bag_kernel, self.bags_values, x)
        return F.matmul(W, self.individuals_mean.to(W.dtype), transa=True)
