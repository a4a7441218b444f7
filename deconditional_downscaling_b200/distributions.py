"""MultivariateNormal with the surface the reference's trainers and likelihoods use
(`.mean`, `.variance`, `.covariance_matrix`, `.lazy_covariance_matrix`,
`.confidence_region()`, `.log_prob()` — SURVEY.md §8b; gpytorch semantics in Appendix A)."""
import math

import torch

from . import functional as F
from .lazy import delazify, lazify


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix, validate_args=False):
        self.loc = mean
        self._covar = lazify(covariance_matrix)

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return delazify(self._covar)

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
        """-0.5 [(v-mu)^T S^{-1} (v-mu) + log det S + k log 2pi] through the libdcgp Cholesky and triangular solve.
        GPU only, like every other computation of the package: a posterior moved with `.cpu()` (the plotting code of the
        runners does that for `.mean` / `.confidence_region()`) raises here instead of factorising on the host."""
        diff = value - self.loc
        S = self.covariance_matrix
        L = F.psd_safe_cholesky(S)
        half = F.trsm(L, diff, trans=False)
        return -0.5 * ((half * half).sum() + F.logdet_from_chol(L) + diff.numel() * math.log(2 * math.pi))
