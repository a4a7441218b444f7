"""gpytorch.models.ExactGP / ApproximateGP surface (SURVEY.md Appendix A rows `ExactGP.__call__`,
`_VariationalStrategy.__call__`)."""
import torch

from ..module import Module


class GP(Module):
    pass


class ExactGPBase(GP):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = None if train_inputs is None else tuple(
            t.unsqueeze(-1) if t.dim() == 1 else t for t in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def _apply(self, fn):
        if self.train_inputs is not None:
            self.train_inputs = tuple(fn(t) for t in self.train_inputs)
            self.train_targets = fn(self.train_targets)
        return super()._apply(fn)


class ApproximateGP(GP):
    def __init__(self, variational_strategy):
        super().__init__()
        self.variational_strategy = variational_strategy

    def __call__(self, inputs, prior=False, **kwargs):
        if inputs.dim() == 1:
            inputs = inputs.unsqueeze(-1)
        return self.variational_strategy(inputs, prior=prior, **kwargs)


class DefaultPredictionStrategy:
    """gpytorch.models.exact_prediction_strategies.DefaultPredictionStrategy as the reference's strategies use it
    (a holder of the training prior, labels and likelihood; cmp_prediction_strategy.py:31-36).  The models of this
    package compute their posteriors themselves (ExactCMP.predict, ExactGP.__call__), so this is only the base
    object a subclass written against gpytorch can build on."""

    def __init__(self, train_inputs, train_prior_dist, train_labels, likelihood, root=None, inv_root=None):
        self.train_inputs, self.train_prior_dist = train_inputs, train_prior_dist
        self.train_labels, self.likelihood = train_labels, likelihood
        self._memoize_cache = {}

    @property
    def lik_train_train_covar(self):
        mvn = self.likelihood(self.train_prior_dist, self.train_inputs)
        return mvn.lazy_covariance_matrix

    @property
    def num_train(self):
        return self.train_labels.numel()
