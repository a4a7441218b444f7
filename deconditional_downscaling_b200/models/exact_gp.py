"""Simplest exact GP regression model (src/models/exact_gp.py:4-23): the mediating regressor
y -> z of the unmatched VBagg data preparation (experiments/swiss_roll/run_experiment.py:129-132)."""
import torch

from .. import functional as F
from ..distributions import MultivariateNormal
from ..kernels import _as_2d
from .base import ExactGPBase


class ExactGP(ExactGPBase):
    def __init__(self, mean_module, covar_module, train_x, train_y, likelihood):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = mean_module
        self.covar_module = covar_module

    def forward(self, x):
        x = _as_2d(x)
        return MultivariateNormal(self.mean_module(x), self.covar_module(x))

    def __call__(self, *args, **kwargs):
        inputs = [_as_2d(i) for i in args]
        if self.training:
            return self.forward(*inputs, **kwargs)
        train_x, xs = self.train_inputs[0], inputs[0]
        kops = self.covar_module.ops(train_x)
        noise = self.likelihood.noise.to(train_x.dtype)
        n = train_x.size(0)
        Ktt = kops.gram(train_x, None) + noise * torch.eye(n, dtype=train_x.dtype, device=train_x.device)
        L = F.psd_safe_cholesky(Ktt)
        Kts = kops.gram(train_x, xs)
        alpha = F.chol_solve(L, (self.train_targets - self.mean_module(train_x)).to(train_x.dtype))
        V = F.trsm(L, Kts, trans=False)
        mean = self.mean_module(xs) + F.matmul(Kts, alpha, transa=True)
        covar = kops.gram(xs, None) - F.matmul(V, V, transa=True)
        return MultivariateNormal(mean, covar)
