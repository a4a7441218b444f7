"""Plain sparse variational GP used by VBagg (src/models/variational_gp.py:7-81)."""
from ..distributions import MultivariateNormal
from ..kernels import _as_2d
from ..variational import CholeskyVariationalDistribution, VariationalStrategy
from .base import ApproximateGP


class VariationalGP(ApproximateGP):
    def __init__(self, inducing_points, mean_module, covar_module):
        variational_strategy = self._set_variational_strategy(inducing_points)
        super().__init__(variational_strategy=variational_strategy)
        self.mean_module = mean_module
        self.covar_module = covar_module

    def _set_variational_strategy(self, inducing_points):
        variational_distribution = CholeskyVariationalDistribution(num_inducing_points=inducing_points.size(0))
        return VariationalStrategy(model=self, inducing_points=inducing_points,
                                   variational_distribution=variational_distribution,
                                   learn_inducing_locations=True)

    def forward(self, inputs):
        inputs = _as_2d(inputs)
        return MultivariateNormal(mean=self.mean_module(inputs), covariance_matrix=self.covar_module(inputs))
