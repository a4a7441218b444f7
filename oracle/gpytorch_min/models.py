"""gpytorch.models restated (test infrastructure).  ExactGP train-mode call =
forward on the stored train inputs (1-D inputs unsqueezed); ApproximateGP call
delegates to the variational strategy; DefaultPredictionStrategy stores the
train prior and `lik_train_train_covar` (SURVEY.md Appendix A rows
`ExactGP.__call__`, `DefaultPredictionStrategy.__init__`).  Sub-classed by the
reference at src/models/exact_cmp.py:7, src/models/variational_cmp.py:9,
src/models/variational_gp.py:7, src/models/exact_gp.py:4,
src/models/prediction_strategies/cmp_prediction_strategy.py:8."""
import sys
import types
import torch
from .module import Module
from .distributions import MultivariateNormal
from . import settings


class GP(Module):
    pass


class ExactGP(GP):
    def __init__(self, train_inputs, train_targets, likelihood):
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        super().__init__()
        if train_inputs is not None:
            self.train_inputs = tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in train_inputs)
            self.train_targets = train_targets
        else:
            self.train_inputs, self.train_targets = None, None
        self.likelihood = likelihood
        self.prediction_strategy = None

    def __call__(self, *args, **kwargs):
        inputs = [i.unsqueeze(-1) if i.dim() == 1 else i for i in args]
        if self.training:
            if settings.debug.on():
                if not all(torch.equal(a, b) for a, b in zip(self.train_inputs, inputs)):
                    raise RuntimeError("You must train on the training inputs!")
            return self.forward(*inputs, **kwargs)
        # eval mode: plain exact GP posterior (only used by the ExactGP
        # mediating regressor, experiments/swiss_roll/core/generation.py:140-204)
        train_x, test_x = self.train_inputs[0], inputs[0]
        prior = self.forward(torch.cat([train_x, test_x], dim=-2))
        n = train_x.size(-2)
        full_mean, full_covar = prior.mean, prior.covariance_matrix
        Ktt = self.likelihood(MultivariateNormal(full_mean[:n], full_covar[:n, :n])).lazy_covariance_matrix
        Kst = full_covar[n:, :n]
        alpha = Ktt.inv_matmul((self.train_targets - full_mean[:n]).unsqueeze(-1)).squeeze(-1)
        mean = full_mean[n:] + Kst @ alpha
        covar = full_covar[n:, n:] - Kst @ Ktt.inv_matmul(Kst.t())
        return MultivariateNormal(mean, covar)


class ApproximateGP(GP):
    def __init__(self, variational_strategy):
        super().__init__()
        self.variational_strategy = variational_strategy

    def __call__(self, inputs, prior=False, **kwargs):
        if inputs.dim() == 1:
            inputs = inputs.unsqueeze(-1)
        return self.variational_strategy(inputs, prior=prior, **kwargs)


class DefaultPredictionStrategy:
    def __init__(self, train_inputs, train_prior_dist, train_labels, likelihood, root=None, inv_root=None):
        self.train_inputs = train_inputs
        self.train_prior_dist = train_prior_dist
        self.train_labels = train_labels.reshape(-1)
        self.likelihood = likelihood
        mvn = self.likelihood(train_prior_dist, train_inputs)
        self.lik_train_train_covar = mvn.lazy_covariance_matrix


exact_prediction_strategies = types.ModuleType(__name__ + ".exact_prediction_strategies")
exact_prediction_strategies.DefaultPredictionStrategy = DefaultPredictionStrategy
sys.modules[exact_prediction_strategies.__name__] = exact_prediction_strategies
