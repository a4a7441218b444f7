"""gpytorch.variational restated (test infrastructure): whitened SVGP with a
Cholesky-parameterised q(u).  Semantics: SURVEY.md Appendix A rows
`VariationalStrategy._cholesky_factor`, `_VariationalStrategy.__call__`,
`CholeskyVariationalDistribution`, `kl_divergence()`.  Sub-classed by the
reference at src/variational/variational_strategy.py:10; used directly at
src/models/variational_gp.py:3-4,34-40."""
import torch
from torch.nn import Parameter
from .module import Module
from .distributions import MultivariateNormal, kl_mvn_mvn
from . import lazy as _lazy
from .utils.cholesky import psd_safe_cholesky
from .utils.memoize import cached, clear_cache_hook
from . import settings


class CholeskyVariationalDistribution(Module):
    def __init__(self, num_inducing_points, batch_shape=torch.Size([]), mean_init_std=1e-3, **kwargs):
        super().__init__()
        self.num_inducing_points = num_inducing_points
        self.mean_init_std = mean_init_std
        self.register_parameter("variational_mean", Parameter(torch.zeros(num_inducing_points)))
        self.register_parameter("chol_variational_covar", Parameter(torch.eye(num_inducing_points)))

    def shape(self):
        return torch.Size([self.num_inducing_points])

    @property
    def dtype(self):
        return self.variational_mean.dtype

    @property
    def device(self):
        return self.variational_mean.device

    def forward(self):
        chol = self.chol_variational_covar
        lower_mask = torch.ones(chol.shape[-2:], dtype=chol.dtype, device=chol.device).tril(0)
        chol = chol.mul(lower_mask)
        return MultivariateNormal(self.variational_mean, _lazy.CholLazyTensor(chol))

    def initialize_variational_distribution(self, prior_dist):
        self.variational_mean.data.copy_(prior_dist.mean)
        self.variational_mean.data.add_(torch.randn_like(prior_dist.mean), alpha=self.mean_init_std)
        self.chol_variational_covar.data.copy_(prior_dist.lazy_covariance_matrix.cholesky().evaluate())


class _VariationalStrategy(Module):
    def __init__(self, model, inducing_points, variational_distribution, learn_inducing_locations=True):
        super().__init__()
        object.__setattr__(self, "model", model)
        inducing_points = inducing_points.clone()
        if inducing_points.dim() == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        if learn_inducing_locations:
            self.register_parameter("inducing_points", Parameter(inducing_points))
        else:
            self.register_buffer("inducing_points", inducing_points)
        self._variational_distribution = variational_distribution
        self.register_buffer("variational_params_initialized", torch.tensor(0))

    def _clear_cache(self):
        clear_cache_hook(self)

    @property
    @cached(name="prior_distribution_memo")
    def prior_distribution(self):
        zeros = torch.zeros(self._variational_distribution.shape(), dtype=self._variational_distribution.dtype,
                            device=self._variational_distribution.device)
        return MultivariateNormal(zeros, _lazy.DiagLazyTensor(torch.ones_like(zeros)))

    @property
    def variational_distribution(self):
        return self._variational_distribution()

    def kl_divergence(self):
        return kl_mvn_mvn(self.variational_distribution, self.prior_distribution)

    def __call__(self, x, prior=False, **kwargs):
        if prior:
            return self.model.forward(x, **kwargs)
        if self.training:
            self._clear_cache()
        if not self.variational_params_initialized.item():
            self._variational_distribution.initialize_variational_distribution(self.prior_distribution)
            self.variational_params_initialized.fill_(1)
        q_u = self.variational_distribution
        return self.forward(x, self.inducing_points, inducing_values=q_u.mean,
                            variational_inducing_covar=q_u.lazy_covariance_matrix, **kwargs)


class VariationalStrategy(_VariationalStrategy):
    def __init__(self, model, inducing_points, variational_distribution, learn_inducing_locations=True):
        super().__init__(model, inducing_points, variational_distribution, learn_inducing_locations)
        self.register_buffer("updated_strategy", torch.tensor(True))

    @cached(name="cholesky_factor", ignore_args=True)
    def _cholesky_factor(self, induc_induc_covar):
        L = psd_safe_cholesky(_lazy.delazify(induc_induc_covar).double())
        return _lazy.TriangularLazyTensor(L)

    def forward(self, x, inducing_points, inducing_values, variational_inducing_covar=None, **kwargs):
        full_inputs = torch.cat([inducing_points, x], dim=-2)
        full_output = self.model.forward(full_inputs, **kwargs)
        full_covar = full_output.lazy_covariance_matrix
        num_induc = inducing_points.size(-2)
        test_mean = full_output.mean[..., num_induc:]
        induc_induc_covar = full_covar[..., :num_induc, :num_induc].add_jitter()
        induc_data_covar = full_covar[..., :num_induc, num_induc:].evaluate()
        data_data_covar = full_covar[..., num_induc:, num_induc:]
        L = self._cholesky_factor(induc_induc_covar)
        interp_term = L.inv_matmul(induc_data_covar.double()).to(full_inputs.dtype)
        predictive_mean = (interp_term.transpose(-1, -2) @ inducing_values.unsqueeze(-1)).squeeze(-1) + test_mean
        middle_term = self.prior_distribution.lazy_covariance_matrix.mul(-1)
        if variational_inducing_covar is not None:
            middle_term = _lazy.SumLazyTensor(variational_inducing_covar, middle_term)
        predictive_covar = _lazy.SumLazyTensor(
            data_data_covar.add_jitter(1e-4),
            _lazy.MatmulLazyTensor(interp_term.transpose(-1, -2), middle_term @ interp_term),
        )
        return MultivariateNormal(predictive_mean, predictive_covar)
