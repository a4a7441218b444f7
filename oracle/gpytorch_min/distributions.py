"""gpytorch.distributions restated (test infrastructure).  MultivariateNormal
semantics from SURVEY.md Appendix A row `MultivariateNormal`; consumed at
src/likelihoods/cmp_likelihood.py:55-58,99, vbagg_gaussian_likelihood.py:74,
experiments/swiss_roll/core/metrics.py:10-27."""
import math
import torch
from torch.distributions import Normal
from . import lazy as _lazy
from .utils.cholesky import psd_safe_cholesky


class base_distributions:
    Normal = Normal


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix, validate_args=False):
        self.loc = mean
        self._covar = _lazy.lazify(covariance_matrix)

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return self._covar.evaluate()

    @property
    def variance(self):
        return self._covar.diag()

    @property
    def stddev(self):
        return self.variance.clamp_min(1e-8).sqrt()

    def confidence_region(self):
        std2 = self.stddev * 2
        return self.mean - std2, self.mean + std2

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    def log_prob(self, value):
        diff = value - self.loc
        inv_quad, logdet = self._covar.inv_quad_logdet(inv_quad_rhs=diff.unsqueeze(-1), logdet=True)
        return -0.5 * (inv_quad + logdet + diff.size(-1) * math.log(2 * math.pi))

    def rsample(self, sample_shape=torch.Size()):
        L = psd_safe_cholesky(self.covariance_matrix)
        eps = torch.randn(*sample_shape, self.loc.numel(), dtype=self.loc.dtype)
        return self.loc + eps @ L.t()

    def __add__(self, other):
        if isinstance(other, MultivariateNormal):
            return MultivariateNormal(self.loc + other.loc, self._covar + other._covar)
        return MultivariateNormal(self.loc + other, self._covar)

    def __mul__(self, c):
        return MultivariateNormal(self.loc * c, self._covar * (c ** 2))


def kl_mvn_mvn(q, p):
    """KL(q || p) through Cholesky factors (exact)."""
    Lq = psd_safe_cholesky(q.covariance_matrix)
    Lp = psd_safe_cholesky(p.covariance_matrix)
    k = q.loc.numel()
    M = torch.linalg.solve_triangular(Lp, Lq, upper=False)
    diff = torch.linalg.solve_triangular(Lp, (p.loc - q.loc).unsqueeze(-1), upper=False)
    logdet_p = 2 * Lp.diagonal().log().sum()
    logdet_q = 2 * Lq.diagonal().abs().log().sum()
    return 0.5 * ((M * M).sum() + (diff * diff).sum() - k + logdet_p - logdet_q)
