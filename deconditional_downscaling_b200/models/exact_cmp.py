"""Exact deconditional GP (src/models/exact_cmp.py:7-169) + its individuals posterior
(src/models/prediction_strategies/cmp_prediction_strategy.py:9-178)."""
import torch

from .. import functional as F
from ..distributions import MultivariateNormal
from ..kernels import _as_2d
from ..lazy import LazyCovariance
from .base import ExactGPBase
from .cmp import CMP
from ..kernels import CMEAggregateKernel
from ..means import CMEAggregateMean


class ExactCMPPredictionStrategy:
    """Caches alpha = (Q + s2 I)^{-1} (z - m) and the Cholesky factor of Q + s2 I
    (cmp_prediction_strategy.py:47-63); predictive mean C alpha + mu*, covariance
    K** - C (Q + s2 I)^{-1} C^T (:101-114, :141-178; the `fast_pred_var` branch is dead code)."""

    def __init__(self, train_individuals, extended_train_bags, train_aggregate_prior_dist,
                 train_aggregate_targets, likelihood):
        self.train_individuals = train_individuals
        self.train_inputs = extended_train_bags
        self.train_prior_dist = train_aggregate_prior_dist
        self.train_labels = train_aggregate_targets
        self.likelihood = likelihood
        mvn = likelihood(train_aggregate_prior_dist, extended_train_bags)
        self.lik_train_train_covar = mvn.lazy_covariance_matrix
        self._Lq = F.psd_safe_cholesky(mvn.covariance_matrix.detach())
        self._alpha = None

    @property
    def posterior_mean_correction_cache(self):
        if self._alpha is None:
            resid = (self.train_labels - self.train_prior_dist.mean).detach().to(self._Lq.dtype)
            self._alpha = F.chol_solve(self._Lq, resid)
        return self._alpha

    def exact_predictive_mean(self, individuals_mean, individuals_to_cme_covar):
        return F.matmul(individuals_to_cme_covar, self.posterior_mean_correction_cache) + individuals_mean

    def exact_predictive_covar(self, individuals_covar, individuals_to_cme_covar):
        C = individuals_to_cme_covar
        V = F.trsm(self._Lq, C.t().contiguous(), trans=False)         # n x N*
        n_star = C.size(0)
        return LazyCovariance(lambda: individuals_covar.evaluate() - F.matmul(V, V, transa=True),
                              lambda: individuals_covar.diag() - (V * V).sum(0), n_star)

    def exact_prediction(self, individuals_mean, individuals_covar, individuals_to_cme_covar):
        return (self.exact_predictive_mean(individuals_mean, individuals_to_cme_covar),
                self.exact_predictive_covar(individuals_covar, individuals_to_cme_covar))


class ExactCMP(ExactGPBase, CMP):
    def __init__(self, train_individuals, extended_train_bags, train_bags, train_aggregate_targets,
                 individuals_mean, individuals_kernel, bag_kernel, lbda, likelihood, use_individuals_noise=True):
        super().__init__(train_inputs=train_bags, train_targets=train_aggregate_targets, likelihood=likelihood)
        self.register_buffer("train_individuals", train_individuals)
        self.register_buffer("train_bags", train_bags)
        self.register_buffer("extended_train_bags", extended_train_bags)
        self.register_buffer("train_aggregate_targets", train_aggregate_targets)
        self.individuals_mean = individuals_mean
        self.individuals_kernel = individuals_kernel
        self.bag_kernel = bag_kernel
        self.noise_kernel = None
        if use_individuals_noise:
            self._init_noise_kernel()
        self.lbda = lbda
        # The reference factors Lambda here on the CPU and moves R in `.to()`; libdcgp has no CPU
        # path, so the (gradient-free) initial factorisation happens at first use on the device
        # (see _ensure_cme_modules), once data and hyper-parameters both live there.  The two CME
        # modules exist from the start - un-factored - so that state_dict() has the reference's
        # `mean_module.*` / `covar_module.*` entries before the first step as well.
        with torch.no_grad():
            mean0 = individuals_mean(_as_2d(train_individuals))
        self.mean_module = CMEAggregateMean(bag_kernel=bag_kernel, bags_values=extended_train_bags,
                                            individuals_mean=mean0, root_inv_bags_covar=None)
        self.covar_module = CMEAggregateKernel(bag_kernel=bag_kernel, bags_values=extended_train_bags,
                                               individuals_covar=None, root_inv_bags_covar=None)
        self.individuals_prediction_strategy = None

    def _ensure_cme_modules(self):
        if self.mean_module.root_inv_bags_covar is None:
            self._init_cmp_mean_covar_modules(self.train_individuals, self.extended_train_bags)

    def _clear_cache(self):
        self.individuals_prediction_strategy = None

    def _init_noise_kernel(self):
        super()._init_noise_kernel(raw_noise=0.)

    def setup_individuals_prediction_strategy(self):
        prior = self.forward(self.train_bags)
        self.individuals_prediction_strategy = ExactCMPPredictionStrategy(
            train_individuals=self.train_individuals, extended_train_bags=self.extended_train_bags,
            train_aggregate_prior_dist=prior, train_aggregate_targets=self.train_aggregate_targets,
            likelihood=self.likelihood)

    def update_cmp_estimate_parameters(self):
        self._ensure_cme_modules()
        return super().update_cmp_estimate_parameters(individuals=self.train_individuals,
                                                      extended_bags_values=self.extended_train_bags)

    def get_individuals_to_cmp_covar(self, input_individuals):
        return super().get_individuals_to_cmp_covar(input_individuals=input_individuals,
                                                    individuals=self.train_individuals,
                                                    bags_values=self.train_bags,
                                                    extended_bags_values=self.extended_train_bags)

    def forward(self, inputs):
        """CME process prior N(m, Q) on bag values (src/models/exact_cmp.py:107-124)."""
        self._ensure_cme_modules()
        self.covar_module.root_inv_bags_covar._cache.clear()
        mean = self.mean_module(inputs)
        covar = self.covar_module(inputs)
        return MultivariateNormal(mean=mean, covariance_matrix=covar)

    def __call__(self, *args, **kwargs):
        inputs = [_as_2d(i) for i in args]
        if self.training:
            ref = _as_2d(self.train_bags)
            if not all(a.shape == ref.shape and torch.equal(a, ref) for a in inputs):
                raise RuntimeError("You must train on the training inputs!")
        return self.forward(*inputs, **kwargs)

    def predict(self, individuals):
        """Posterior on individuals (src/models/exact_cmp.py:126-155)."""
        if self.individuals_prediction_strategy is None:
            self.setup_individuals_prediction_strategy()
        individuals = _as_2d(individuals)
        individuals_mean = self.individuals_mean(individuals)
        individuals_covar = self.individuals_kernel(individuals)
        C = self.get_individuals_to_cmp_covar(individuals)
        strat = self.individuals_prediction_strategy
        mean = strat.exact_predictive_mean(individuals_mean=individuals_mean, individuals_to_cme_covar=C)
        covar = strat.exact_predictive_covar(individuals_covar=individuals_covar, individuals_to_cme_covar=C)
        return MultivariateNormal(mean=mean, covariance_matrix=covar)

    def _apply(self, fn):
        out = super()._apply(fn)
        # device / dtype changed: Lambda^{-1} is re-factored lazily where the data now lives
        self.mean_module.root_inv_bags_covar = None
        self.covar_module.root_inv_bags_covar = None
        self.covar_module.individuals_covar = None
        self.individuals_prediction_strategy = None
        return out

    def train(self, mode=True):
        return super().train(mode)
