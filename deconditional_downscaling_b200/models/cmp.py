"""CMP mixin: what ExactCMP and VariationalCMP share (src/models/cmp.py:8-104)."""
import torch

from .. import functional as F
from ..kernels import CMEAggregateKernel, DeltaKernel, KernelInverse, KernelMatrix, ScaleKernel, _as_2d
from ..means import CMEAggregateMean


class CMP:
    def _init_noise_kernel(self, raw_noise):
        """sigma_ind^2 * delta kernel, raw (pre-softplus) init `raw_noise` (src/models/cmp.py:11-20)."""
        self.noise_kernel = ScaleKernel(base_kernel=DeltaKernel())
        self.noise_kernel.initialize(raw_outputscale=raw_noise * torch.ones(1))

    def _individuals_covar(self, individuals):
        """k(x, x) [+ sigma_ind^2 I], kept symbolic (src/models/cmp.py:60-62)."""
        K = self.individuals_kernel(individuals)
        if self.noise_kernel is not None:
            K = K + self.noise_kernel(individuals)
        return K

    def _get_cmp_estimate_parameters(self, individuals, extended_bags_values, detach=False):
        """mu(x), K (symbolic), and the inverse of Lambda = l(y~, y~) + lbda N I
        (src/models/cmp.py:48-68; the reference returns the root R of Lambda^{-1})."""
        individuals = _as_2d(individuals)
        ext = _as_2d(extended_bags_values)
        mean = self.individuals_mean(individuals)
        covar = self._individuals_covar(individuals)
        N = ext.size(0)
        inv = KernelInverse(self.bag_kernel.ops(ext), ext, self.lbda * N, detach=detach)
        return (mean.detach() if detach else mean), covar, inv

    def _init_cmp_mean_covar_modules(self, individuals, extended_bags_values):
        """Builds CMEAggregateMean / CMEAggregateKernel with Lambda^{-1} evaluated WITHOUT
        gradient, like the reference's `torch.no_grad()` block (src/models/cmp.py:22-46), while K
        stays symbolic so its hyper-parameters get gradient from the first epoch on."""
        with torch.no_grad():
            mean, _, inv = self._get_cmp_estimate_parameters(individuals, extended_bags_values, detach=True)
        covar = self._individuals_covar(_as_2d(individuals))
        self.mean_module = CMEAggregateMean(bag_kernel=self.bag_kernel, bags_values=extended_bags_values,
                                            individuals_mean=mean, root_inv_bags_covar=inv)
        self.covar_module = CMEAggregateKernel(bag_kernel=self.bag_kernel, bags_values=extended_bags_values,
                                               individuals_covar=covar, root_inv_bags_covar=inv)

    def update_cmp_estimate_parameters(self, individuals, extended_bags_values):
        """Refresh mu, K and Lambda^{-1} with gradient enabled (src/models/cmp.py:70-85): from now on
        the bag kernel receives gradient through Lambda^{-1} as well as through l(y~, y)."""
        mean, covar, inv = self._get_cmp_estimate_parameters(individuals, extended_bags_values)
        self.mean_module.individuals_mean = mean
        self.mean_module.bags_values = extended_bags_values
        self.mean_module.root_inv_bags_covar = inv
        self.covar_module.bags_values = extended_bags_values
        self.covar_module.individuals_covar = covar
        self.covar_module.root_inv_bags_covar = inv

    def get_individuals_to_cmp_covar(self, input_individuals, individuals, bags_values, extended_bags_values):
        """C = k(x*, x) Lambda^{-1} l(y~, y)   (N* x n)  (src/models/cmp.py:87-104)."""
        W = self.covar_module.weights(bags_values)
        return self.individuals_kernel(_as_2d(input_individuals), _as_2d(individuals)).matmul(W)
