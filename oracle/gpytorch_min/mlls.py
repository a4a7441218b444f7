"""gpytorch.mlls restated (test infrastructure).  ExactMarginalLogLikelihood =
marginal log_prob / n; used at
experiments/swiss_roll/models/exact_cmp.py:81,97.  MarginalLogLikelihood is
the base of src/mlls/bag_variational_elbo.py:5."""
import sys
import types
from .module import Module


class MarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model


class ExactMarginalLogLikelihood(MarginalLogLikelihood):
    def forward(self, function_dist, target, *params):
        output = self.likelihood(function_dist, *params)
        res = output.log_prob(target)
        return res / target.size(-1)


class VariationalELBO(MarginalLogLikelihood):
    def __init__(self, likelihood, model, num_data, beta=1.0, combine_terms=True):
        super().__init__(likelihood, model)
        self.num_data, self.beta = num_data, beta

    def forward(self, variational_dist_f, target, **kwargs):
        ll = self.likelihood.expected_log_prob(target, variational_dist_f, **kwargs).sum(-1).div(target.size(0))
        kl = self.model.variational_strategy.kl_divergence().div(self.num_data / self.beta)
        return ll - kl


# `from gpytorch.mlls.marginal_log_likelihood import MarginalLogLikelihood`
marginal_log_likelihood = types.ModuleType(__name__ + ".marginal_log_likelihood")
marginal_log_likelihood.MarginalLogLikelihood = MarginalLogLikelihood
sys.modules[marginal_log_likelihood.__name__] = marginal_log_likelihood
