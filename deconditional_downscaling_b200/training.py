"""Training-step drivers: one process per GPU, data-parallel over bag minibatches.

`ElboTrainer.step` is the hot loop body of the reference's variational trainers
(experiments/downscaling/models/variational_cmp.py:163-186,
experiments/swiss_roll/models/vbagg.py:88-101): q = model(x) -> ELBO -> backward -> Adam,
plus - when torch.distributed is initialised - ONE NCCL all-reduce of the flat gradient
buffer (every rank draws its own minibatch; gradients are averaged).
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import functional as F
from .likelihoods import CMPLikelihood, VBaggGaussianLikelihood
from .mlls import BagVariationalELBO


class FlatGrads:
    """All parameter gradients as views into one contiguous buffer, so the data-parallel
    exchange is a single collective (payload ~ M^2 + M d + M + O(10) values)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dtype = self.params[0].dtype
        self.flat = torch.zeros(total, dtype=dtype, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()


class ElboTrainer:
    """VariationalCMP (+CMPLikelihood) or VariationalGP (+VBaggGaussianLikelihood) ELBO steps."""

    def __init__(self, model, likelihood, num_data, beta=1.0, lr=0.01):
        self.model, self.likelihood = model, likelihood
        self.elbo = BagVariationalELBO(likelihood, model, num_data=num_data, beta=beta)
        self.params = list(model.parameters()) + list(likelihood.parameters())
        # q(u) must exist before the optimizer / flat buffer are laid out
        model.variational_strategy.ensure_initialized()
        self.grads = FlatGrads(self.params)
        self.opt = torch.optim.Adam(self.params, lr=lr)
        model.train()
        likelihood.train()

    def loss(self, x, z, **kw):
        q = self.model(x)
        if isinstance(self.likelihood, CMPLikelihood):
            kw = self.model.get_elbo_computation_parameters(bags_values=kw["bags_values"],
                                                            extended_bags_values=kw["extended_bags_values"])
        return -self.elbo(variational_dist_f=q, target=z, **kw)

    def step(self, x, z, **kw):
        """One ELBO step on device-resident tensors; returns the (device) loss."""
        self.grads.zero()
        try:
            # optimistic pass: no host read-back between the factorisations, ONE check before the update
            with F.deferred_pd_checks() as pending:
                loss = self.loss(x, z, **kw)
                loss.backward()
            pending.check()
        except F.NotPSDError:
            # psd_safe_cholesky semantics (jitter retries, then NotPSDError) with immediate checks
            self.grads.zero()
            loss = self.loss(x, z, **kw)
            loss.backward()
        self.grads.allreduce_mean()
        self.opt.step()
        return loss.detach()

    def step_from_host(self, host_batch: Dict[str, torch.Tensor], device):
        """End-to-end step: pinned host tensors -> device, step, loss back to the host."""
        # persistent device staging buffers (one per input name and shape): no allocator traffic per step
        stage = self.__dict__.setdefault("_stage", {})
        dev = {}
        for k, v in host_batch.items():
            if not torch.is_tensor(v):
                dev[k] = v
                continue
            key = (k, tuple(v.shape), v.dtype)
            buf = stage.get(key)
            if buf is None:
                buf = stage[key] = torch.empty(v.shape, dtype=v.dtype, device=device)
            buf.copy_(v, non_blocking=True)
            dev[k] = buf
        x, z = dev.pop("x"), dev.pop("z")
        return float(self.step(x, z, **dev).item())
