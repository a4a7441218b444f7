"""Aggregate-observation GP baseline of the swiss-roll experiment (src/models/bagged_gp.py:9-194 and
src/models/prediction_strategies/bagged_gp_prediction_strategy.py:8-170).

The bags are assumed independent: the prior on the aggregate values is N(mean_a(mu), diag(1_a^T K_aa 1_a / n_a^2))
and an individual only co-varies with the aggregate of its own bag.  The reference extracts the diagonal blocks
of the N x N kernel matrix in python loops over the bags; here the same quantities come from the segmented
Gram reductions of the hot path (dcgp_gram_bag_blocksum, dcgp_gram_bagsum), so the N x N matrix is never built
for training."""
import torch

from .. import functional as F
from ..distributions import MultivariateNormal
from ..kernels import _as_2d
from ..lazy import DiagLazyTensor
from ..likelihoods import bag_offsets
from .base import ExactGPBase


class BaggedGP(ExactGPBase):
    def __init__(self, train_individuals, bags_sizes, train_aggregate_targets, individuals_mean, individuals_kernel, likelihood):
        sizes = torch.as_tensor(bags_sizes, dtype=torch.int32).reshape(-1)
        super().__init__(train_inputs=sizes.clone(), train_targets=train_aggregate_targets, likelihood=likelihood)
        self.register_buffer("bags_sizes", sizes)
        self.register_buffer("train_individuals", _as_2d(train_individuals))
        self.register_buffer("train_aggregate_targets", train_aggregate_targets)
        self.individuals_mean = individuals_mean
        self.individuals_kernel = individuals_kernel
        self._cache = None

    def _apply(self, fn):
        self._cache = None
        return super()._apply(fn)

    def _clear_cache(self):
        self._cache = None

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, inputs, bags_sizes):
        """Prior on the aggregate values of the given bags (bagged_gp.py:81-104)."""
        inputs = _as_2d(inputs)
        sizes_list = [int(s) for s in torch.as_tensor(bags_sizes).reshape(-1).tolist()]
        offsets = bag_offsets(sizes_list, inputs.device)
        sizes = torch.tensor(sizes_list, dtype=inputs.dtype, device=inputs.device)
        mean = F.bag_sum(self.individuals_mean(inputs).to(inputs.dtype), offsets) / sizes
        var = self.individuals_kernel.ops(inputs).bag_blocksum(inputs, offsets) / (sizes * sizes)
        return MultivariateNormal(mean, DiagLazyTensor(var))

    # ---- prediction ------------------------------------------------------------------------------------
    def _train_cache(self):
        if self._cache is None:
            x, sizes_list = self.train_individuals, self.bags_sizes.tolist()
            prior = self.forward(x, sizes_list)
            sizes = torch.tensor(sizes_list, dtype=x.dtype, device=x.device)
            total = prior.variance + self.likelihood.noise.to(x.dtype) / sizes          # diagonal of the aggregate covariance
            alpha = (self.train_aggregate_targets.to(x.dtype) - prior.mean) / total
            if not torch.is_grad_enabled():
                total, alpha = total.detach(), alpha.detach()
            self._cache = (total, alpha)
        return self._cache

    def get_individuals_to_aggregate_covar(self, input_individuals, input_bags_sizes):
        """(N*, n) matrix whose column a holds, for the inputs that ARE training bag a, the mean covariance with
        that bag, and zeros elsewhere (bagged_gp.py:115-150)."""
        xs, xt = _as_2d(input_individuals), self.train_individuals
        t_sizes = self.bags_sizes.tolist()
        t_off = [0]
        for s in t_sizes:
            t_off.append(t_off[-1] + s)
        offsets = bag_offsets(t_sizes, xt.device)
        # column sums over every training bag, then keep only the (input bag == training bag) blocks
        sums = self.individuals_kernel.ops(xt).bagsum(xs, xt, offsets)                   # n x N*
        sizes = torch.tensor(t_sizes, dtype=xs.dtype, device=xs.device)
        cross = (sums / sizes.unsqueeze(1)).t()                                           # N* x n
        mask = torch.zeros_like(cross)
        if xs.size() == xt.size() and torch.allclose(xs, xt):
            for a in range(len(t_sizes)):
                mask[t_off[a]:t_off[a + 1], a] = 1
        else:
            i_sizes = [int(s) for s in torch.as_tensor(input_bags_sizes).reshape(-1).tolist()]
            i_off = [0]
            for s in i_sizes:
                i_off.append(i_off[-1] + s)
            for a in range(len(t_sizes)):
                for b in range(len(i_sizes)):
                    if t_sizes[a] == i_sizes[b] and torch.allclose(xt[t_off[a]:t_off[a + 1]], xs[i_off[b]:i_off[b + 1]]):
                        mask[i_off[b]:i_off[b + 1], a] = 1
        return cross * mask

    def predict(self, individuals, input_bags_sizes):
        """Posterior on the individuals given the aggregate observations (bagged_gp.py:152-180)."""
        xs = _as_2d(individuals)
        total, alpha = self._train_cache()
        C = self.get_individuals_to_aggregate_covar(xs, input_bags_sizes)
        mean = self.individuals_mean(xs).to(xs.dtype) + F.matmul(C, alpha)
        Cs = C / total.sqrt().unsqueeze(0)
        covar = self.individuals_kernel.ops(xs).gram(xs, None) - F.matmul(Cs, Cs, transb=True)
        return MultivariateNormal(mean, covar)
