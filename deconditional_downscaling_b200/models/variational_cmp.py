"""Sparse variational conditional mean process (src/models/variational_cmp.py:9-113)."""
import torch

from ..distributions import MultivariateNormal
from ..kernels import KernelInverse, _as_2d
from ..variational import CholeskyVariationalDistribution, VariationalStrategy
from .base import ApproximateGP
from .cmp import CMP


class VariationalCMP(ApproximateGP, CMP):
    def __init__(self, inducing_points, individuals_mean, individuals_kernel, bag_kernel, lbda,
                 use_individuals_noise=True):
        variational_strategy = self._set_variational_strategy(inducing_points)
        super().__init__(variational_strategy=variational_strategy)
        self.individuals_mean = individuals_mean
        self.individuals_kernel = individuals_kernel
        self.bag_kernel = bag_kernel
        self.noise_kernel = None
        self.lbda = lbda
        if use_individuals_noise:
            self._init_noise_kernel()

    def _set_variational_strategy(self, inducing_points):
        variational_distribution = CholeskyVariationalDistribution(num_inducing_points=inducing_points.size(0))
        return VariationalStrategy(model=self, inducing_points=inducing_points,
                                   variational_distribution=variational_distribution,
                                   learn_inducing_locations=True)

    def _init_noise_kernel(self):
        super()._init_noise_kernel(raw_noise=0.)

    def forward(self, inputs):
        """Prior N(mu(x), k(x, x)) (src/models/variational_cmp.py:67-84)."""
        inputs = _as_2d(inputs)
        return MultivariateNormal(mean=self.individuals_mean(inputs), covariance_matrix=self.individuals_kernel(inputs))

    def get_elbo_computation_parameters(self, bags_values, extended_bags_values):
        """Lambda^{-1} for Lambda = l(y~, y~) + lbda N I (returned under the reference's key
        `root_inv_extended_bags_covar`), l(y, y~) kept symbolic, and sigma_ind^2
        (src/models/variational_cmp.py:86-113)."""
        ext = _as_2d(extended_bags_values)
        N = ext.size(0)
        inv = KernelInverse(self.bag_kernel.ops(ext), ext, self.lbda * N)
        output = {"root_inv_extended_bags_covar": inv,
                  "bags_to_extended_bags_covar": self.bag_kernel(_as_2d(bags_values), ext)}
        if self.noise_kernel is not None:
            output.update({"individuals_noise": self.noise_kernel.outputscale})
        return output
