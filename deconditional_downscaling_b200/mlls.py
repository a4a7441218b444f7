"""Objective modules: gpytorch's ExactMarginalLogLikelihood (as used at
experiments/swiss_roll/models/exact_cmp.py:81,97) and the reference's BagVariationalELBO
(src/mlls/bag_variational_elbo.py:5-69)."""
import torch

from .module import Module


class MarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model


class ExactMarginalLogLikelihood(MarginalLogLikelihood):
    """log N(target | m, Q + sigma^2 I) / n  (SURVEY.md Appendix A)."""

    def forward(self, function_dist, target, *params):
        output = self.likelihood(function_dist, *params)
        return output.log_prob(target) / target.size(-1)


class BagVariationalELBO(MarginalLogLikelihood):
    """ELL / target.size(0) - KL / (num_data / beta) + log_prior / num_data - added_loss
    (src/mlls/bag_variational_elbo.py:47-69); no priors / added losses exist on the hot path."""

    def __init__(self, likelihood, model, num_data, beta=1.0, combine_terms=True):
        super().__init__(likelihood, model)
        self.combine_terms = combine_terms
        self.num_data = num_data
        self.beta = beta

    def _log_likelihood_term(self, variational_dist_f, target, **kwargs):
        return self.likelihood.expected_log_prob(target, variational_dist_f, **kwargs).sum(-1)

    def forward(self, variational_dist_f, target, **kwargs):
        log_likelihood = self._log_likelihood_term(variational_dist_f, target, **kwargs).div(target.size(0))
        kl_divergence = self.model.variational_strategy.kl_divergence().div(self.num_data / self.beta)
        log_prior = torch.zeros_like(log_likelihood)
        if self.combine_terms:
            return log_likelihood - kl_divergence + log_prior
        return log_likelihood, kl_divergence, log_prior


class VariationalELBO(BagVariationalELBO):
    """gpytorch.mlls.VariationalELBO (used by the krigging baseline,
    experiments/downscaling/models/krigging.py:173): same combination rule."""
