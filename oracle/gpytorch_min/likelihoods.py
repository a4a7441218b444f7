"""gpytorch.likelihoods restated (test infrastructure).  GaussianLikelihood:
noise = softplus(raw_noise) + 1e-4 (GreaterThan(1e-4)), state-dict key
`noise_covar.raw_noise`; `marginal` adds noise*I; calling the likelihood on a
MultivariateNormal returns the marginal (SURVEY.md Appendix A rows
`Parameter transforms`, `GaussianLikelihood.marginal`).  Sub-classed by the
reference at src/likelihoods/cmp_likelihood.py:7 and
src/likelihoods/vbagg_gaussian_likelihood.py:7."""
import math
import torch
from torch.nn import Parameter
from .module import Module
from .constraints import GreaterThan
from .distributions import MultivariateNormal, base_distributions
from . import lazy as _lazy


class Likelihood(Module):
    def __call__(self, input, *args, **kwargs):
        if torch.is_tensor(input):
            return self.forward(input, *args, **kwargs)
        if isinstance(input, MultivariateNormal):
            return self.marginal(input, *args, **kwargs)
        raise RuntimeError("Likelihoods expects a MultivariateNormal or a Tensor input")


class HomoskedasticNoise(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size()):
        super().__init__()
        if noise_constraint is None:
            noise_constraint = GreaterThan(1e-4)
        self.register_parameter("raw_noise", Parameter(torch.zeros(*batch_shape, 1)))
        self.register_constraint("raw_noise", noise_constraint)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value).to(self.raw_noise)
        self.initialize(raw_noise=self.raw_noise_constraint.inverse_transform(value))

    def forward(self, *params, shape=None, **kwargs):
        noise = self.noise
        if shape is None:
            p = params[0] if torch.is_tensor(params[0]) else params[0][0]
            shape = p.shape if len(p.shape) == 1 else p.shape[:-1]
        n = shape[-1]
        return _lazy.DiagLazyTensor(noise.expand(n))


class GaussianLikelihood(Likelihood):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise(noise_prior=noise_prior, noise_constraint=noise_constraint,
                                              batch_shape=batch_shape)

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.initialize(noise=value)

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    def _shaped_noise_covar(self, base_shape, *params, **kwargs):
        return self.noise_covar(*params, shape=base_shape, **kwargs)

    def forward(self, function_samples, *params, **kwargs):
        noise = self._shaped_noise_covar(function_samples.shape, *params, **kwargs).diag()
        return base_distributions.Normal(function_samples, noise.sqrt())

    def marginal(self, function_dist, *params, **kwargs):
        mean, covar = function_dist.mean, function_dist.lazy_covariance_matrix
        noise_covar = self._shaped_noise_covar(mean.shape, *params, **kwargs)
        return MultivariateNormal(mean, covar + noise_covar)

    def expected_log_prob(self, target, input, *params, **kwargs):
        mean, variance = input.mean, input.variance
        noise = self._shaped_noise_covar(mean.shape, *params, **kwargs).diag()
        res = ((target - mean) ** 2 + variance) / noise + noise.log() + math.log(2 * math.pi)
        return res.mul(-0.5)
