"""Likelihoods: GaussianLikelihood (gpytorch semantics, SURVEY.md Appendix A: noise =
softplus(raw_noise) + 1e-4, state-dict key `noise_covar.raw_noise`) and the reference's two
bag likelihoods, src/likelihoods/cmp_likelihood.py:7-116 and
src/likelihoods/vbagg_gaussian_likelihood.py:7-111, evaluated in their minimal algebraic
form on libdcgp kernels."""
import math

import numpy as np
import torch
from torch import nn
from torch.nn.functional import softplus

from . import functional as F
from .distributions import MultivariateNormal
from .lazy import DenseMatrix, delazify
from .constraints import GreaterThan
from .module import Module


class HomoskedasticNoise(Module):
    def __init__(self):
        super().__init__()
        self.register_parameter("raw_noise", nn.Parameter(torch.zeros(1)))
        self.raw_noise_constraint = GreaterThan(1e-4)

    @property
    def noise(self):
        return softplus(self.raw_noise) + 1e-4

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value).to(self.raw_noise) - 1e-4
        self.initialize(raw_noise=value + torch.log(-torch.expm1(-value)))

    def forward(self, *params, shape=None, **kwargs):
        n = shape[-1] if shape is not None else params[0].shape[0]
        return DenseMatrix(torch.diag_embed(self.noise.expand(n)))


class Likelihood(Module):
    """gpytorch.likelihoods.Likelihood base."""


class GaussianLikelihood(Likelihood):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=None, **kwargs):
        super().__init__()
        self.noise_covar = HomoskedasticNoise()

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    @property
    def raw_noise(self):
        return self.noise_covar.raw_noise

    def marginal(self, function_dist, *params, **kwargs):
        mean, covar = function_dist.mean, function_dist.covariance_matrix
        noise = self.noise.to(covar.dtype)
        eye = torch.eye(mean.numel(), dtype=covar.dtype, device=covar.device)
        return MultivariateNormal(mean, covar + noise * eye)

    def forward(self, function_samples, *params, **kwargs):
        return torch.distributions.Normal(function_samples, self.noise.to(function_samples.dtype).sqrt())

    def __call__(self, input, *args, **kwargs):
        if torch.is_tensor(input):
            return self.forward(input, *args, **kwargs)
        return self.marginal(input, *args, **kwargs)

    def expected_log_prob(self, target, input, *params, **kwargs):
        mean, variance = input.mean, input.variance
        noise = self.noise.to(mean.dtype)
        return -0.5 * (((target - mean) ** 2 + variance) / noise + noise.log() + math.log(2 * math.pi))


def _aggregation_operator(root_inv_extended_bags_covar, bags_to_extended_bags_covar):
    """A = Lambda^{-1} l(y~, y)  (N x n): the deconditioning operator.  The reference forms it as
    R (B R)^T with an explicit root R (cmp_likelihood.py:61, 107)."""
    B = bags_to_extended_bags_covar
    Bt = B.t().evaluate() if hasattr(B, "evaluate") else B.t().contiguous()
    return root_inv_extended_bags_covar.solve(Bt)


class CMPLikelihood(GaussianLikelihood):
    def __init__(self, use_individuals_noise=True, **kwargs):
        super().__init__(**kwargs)
        self.use_individuals_noise = use_individuals_noise

    def expected_log_prob(self, observations, variational_dist, root_inv_extended_bags_covar,
                          bags_to_extended_bags_covar, individuals_noise=None):
        if self.use_individuals_noise:
            return self._compute_expected_log_prob_with_individuals_noise(
                observations, variational_dist, root_inv_extended_bags_covar, bags_to_extended_bags_covar,
                individuals_noise)
        return self._compute_expected_log_prob_without_individuals_noise(
            observations, variational_dist, root_inv_extended_bags_covar, bags_to_extended_bags_covar)

    def _compute_expected_log_prob_with_individuals_noise(self, observations, variational_dist,
                                                          root_inv_extended_bags_covar,
                                                          bags_to_extended_bags_covar, individuals_noise):
        """-0.5 (n log 2pi + log|C| + e^T C^{-1} e + tr(C^{-1} A^T Sigma_q A)),
        C = sigma_ind^2 A^T A + sigma_agg^2 I, e = z - A^T mu_q  (cmp_likelihood.py:41-78).  The
        reference's chol(Sigma_q) (:58) is not needed: G = A^T Sigma_q A is formed directly."""
        z = observations
        n = z.numel()
        A = _aggregation_operator(root_inv_extended_bags_covar, bags_to_extended_bags_covar)
        H = F.matmul(A, A, transa=True)
        a = F.matmul(A, variational_dist.mean, transa=True)
        G = variational_dist.quad_form(A) if hasattr(variational_dist, "quad_form") else \
            F.matmul(A, F.matmul(variational_dist.covariance_matrix, A), transa=True)
        noise = self.noise.to(z.dtype)
        eye = torch.eye(n, dtype=z.dtype, device=z.device)
        C = individuals_noise.to(z.dtype) * H + noise * eye
        logdet, mean_term, covar_term = F.gaussian_terms(C, z - a, G)
        return -0.5 * (n * math.log(2 * math.pi) + logdet + mean_term + covar_term)

    def _compute_expected_log_prob_without_individuals_noise(self, observations, variational_dist,
                                                             root_inv_extended_bags_covar,
                                                             bags_to_extended_bags_covar):
        """-0.5 (n log(2 pi s2) + (|z - A^T mu_q|^2 + sum_i v_i |A_i.|^2) / s2) using only
        diag(Sigma_q) (cmp_likelihood.py:80-116, :99)."""
        z = observations
        n = z.numel()
        A = _aggregation_operator(root_inv_extended_bags_covar, bags_to_extended_bags_covar)
        err = z - F.matmul(A, variational_dist.mean, transa=True)
        covar_term = ((A * A).sum(-1) * variational_dist.variance).sum()
        noise = self.noise.to(z.dtype)
        return -0.5 * (n * torch.log(2 * math.pi * noise) + ((err * err).sum() + covar_term) / noise)


_OFFSETS = {}


def bag_offsets(bags_sizes, device):
    """int64 device array of n_bags + 1 row offsets of contiguous bags.  Lists are memoised (a trainer passes the same
    sizes every step; the upload is a pageable host-to-device copy, which also has no place inside a CUDA graph)."""
    if torch.is_tensor(bags_sizes):
        sizes = bags_sizes.to(dtype=torch.int64, device="cpu")
        key = None
    else:
        key = (tuple(int(s) for s in bags_sizes), str(device))
        hit = _OFFSETS.get(key)
        if hit is not None:
            return hit
        sizes = torch.tensor(key[0], dtype=torch.int64)
    out = torch.cat([torch.zeros(1, dtype=torch.int64), sizes.cumsum(0)]).to(device)
    if key is not None:
        if len(_OFFSETS) > 64:
            _OFFSETS.clear()
        _OFFSETS[key] = out
    return out


class VBaggGaussianLikelihood(GaussianLikelihood):
    """y_a = mean(f(B_a)) + eps, eps ~ N(0, sigma^2 / n_a)
    (src/likelihoods/vbagg_gaussian_likelihood.py:7-111)."""

    def marginal(self, function_dist, bags_sizes):
        mean, covar = function_dist.mean, function_dist.covariance_matrix
        sizes = torch.as_tensor(bags_sizes, dtype=mean.dtype, device=mean.device)
        return MultivariateNormal(mean, covar + torch.diag_embed(self.noise.to(mean.dtype) / sizes))

    def forward(self, individuals_evaluations, bags_sizes):
        offsets = bag_offsets(bags_sizes, individuals_evaluations.device)
        sizes = torch.as_tensor(bags_sizes, dtype=individuals_evaluations.dtype, device=individuals_evaluations.device)
        agg = F.bag_sum(individuals_evaluations, offsets) / sizes
        # the reference passes the variance vector as `scale` (vbagg_gaussian_likelihood.py:47-50)
        return torch.distributions.Normal(loc=agg, scale=self.noise.to(agg.dtype) / sizes)

    def expected_log_prob(self, observations, variational_dist, bags_sizes):
        """-0.5 sum_a [(z_a^2 - 2 z_a S_a / n_a + (T_a + S_a^2) / n_a^2) n_a / s2 + log(2 pi s2 / n_a)],
        S_a = 1_a^T mu_q, T_a = 1_a^T Sigma_q 1_a  (vbagg_gaussian_likelihood.py:78-111).  The dense
        N x N covariance of :74 and the python loops of :75,:99-101 are replaced by the fused
        bag-aggregation kernels (the posterior hands out S_a and T_a directly)."""
        z = observations
        offsets = bag_offsets(bags_sizes, z.device)
        if hasattr(variational_dist, "bag_stats"):
            S, T = variational_dist.bag_stats(offsets)
        else:
            S = F.bag_sum(variational_dist.mean, offsets)
            rows = F.bag_sum(variational_dist.covariance_matrix, offsets)          # n_bags x N
            T = F.bag_sum(rows.t().contiguous(), offsets).diagonal()
        return F.vbagg_ell(z, S, T, offsets, self.noise.to(z.dtype))
