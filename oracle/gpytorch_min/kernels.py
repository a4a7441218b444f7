"""gpytorch.kernels restated (test infrastructure).  Behaviour per
gpytorch==1.4.0 as summarised in SURVEY.md Appendix A rows `Kernel.__call__`,
`Squared distance`, `RBF`, `Matérn ν=1.5`, `gpytorch RFFKernel`:

* `Kernel.__call__`: select `active_dims`, unsqueeze 1-D inputs, `k(x)` means
  `k(x, x)`; result is a (here: eager) lazy tensor.
* squared distance by the mean-centred quadratic expansion
  `[-2 x1, |x1|^2, 1] @ [x2, 1, |x2|^2]^T`, zero diagonal when x1 equals x2 and
  no input requires grad, `clamp_min(0)`; distance = sqrt(clamp_min(sq, 1e-30)).
* RBF = exp(-sq_dist(x1/l, x2/l)/2); Matern: inputs centred by mean(x1) then
  scaled; ScaleKernel multiplies by softplus(raw_outputscale) and inherits the
  base kernel's active_dims; `k1 + k2` -> AdditiveKernel exposing `.kernels`.

Called from the reference at src/models/cmp.py:60-65,98-99,
src/models/variational_cmp.py:79,99,103, src/kernels/cme_aggregate_kernel.py:38-39,
src/means/cme_aggregate_mean.py:36, src/kernels/rff_kernel.py:32 (subclass).
"""
import math
import torch
from torch.nn import ModuleList, Parameter
from torch.nn.functional import softplus
from .module import Module
from .constraints import Positive
from . import lazy as _lazy


def sq_dist(x1, x2, x1_eq_x2=False):
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    same = x1_eq_x2 and not x1.requires_grad and not x2.requires_grad
    if same:
        x2_norm, x2_pad = x1_norm, x1_pad
    else:
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if same:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min(0)


def dist(x1, x2, x1_eq_x2=False):
    return sq_dist(x1, x2, x1_eq_x2).clamp_min(1e-30).sqrt()


class Kernel(Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None,
                 lengthscale_prior=None, lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        self._batch_shape = batch_shape
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(active_dims, dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self.eps = eps
        if self.has_lengthscale:
            nd = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale", Parameter(torch.zeros(*batch_shape, 1, nd)))
            self.register_constraint("raw_lengthscale", Positive())

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def lengthscale(self):
        if self.has_lengthscale:
            return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)
        return None

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value).to(self.raw_lengthscale)
        self.initialize(raw_lengthscale=self.raw_lengthscale_constraint.inverse_transform(value))

    def covar_dist(self, x1, x2, diag=False, last_dim_is_batch=False, square_dist=False,
                   dist_postprocess_func=None, postprocess=True, **params):
        x1_eq_x2 = torch.equal(x1, x2)
        if diag:
            if x1_eq_x2:
                res = torch.zeros(x1.shape[:-1], dtype=x1.dtype, device=x1.device)
            else:
                res = (x1 - x2).pow(2).sum(-1)
                if not square_dist:
                    res = res.clamp_min(1e-30).sqrt()
        else:
            res = sq_dist(x1, x2, x1_eq_x2) if square_dist else dist(x1, x2, x1_eq_x2)
        if postprocess and dist_postprocess_func is not None:
            res = dist_postprocess_func(res)
        return res

    def forward(self, x1, x2, diag=False, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        x1_, x2_ = x1, x2
        if self.active_dims is not None:
            x1_ = x1_.index_select(-1, self.active_dims)
            if x2_ is not None:
                x2_ = x2_.index_select(-1, self.active_dims)
        if x1_.dim() == 1:
            x1_ = x1_.unsqueeze(1)
        if x2_ is not None and x2_.dim() == 1:
            x2_ = x2_.unsqueeze(1)
        if x2_ is None:
            x2_ = x1_
        if diag:
            return _lazy.delazify(self.forward(x1_, x2_, diag=True, **params))
        return _lazy.LazyEvaluatedKernelTensor(x1_, x2_, self, **params)

    def __add__(self, other):
        kernels = []
        kernels += self.kernels if isinstance(self, AdditiveKernel) else [self]
        kernels += other.kernels if isinstance(other, AdditiveKernel) else [other]
        return AdditiveKernel(*kernels)


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        x1_ = x1.div(self.lengthscale)
        x2_ = x2.div(self.lengthscale)
        return self.covar_dist(x1_, x2_, square_dist=True, diag=diag,
                               dist_postprocess_func=lambda d: d.div(-2).exp(), postprocess=True)


class MaternKernel(Kernel):
    has_lengthscale = True

    def __init__(self, nu=2.5, **kwargs):
        if nu not in {0.5, 1.5, 2.5}:
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super().__init__(**kwargs)
        self.nu = nu

    def forward(self, x1, x2, diag=False, **params):
        mean = x1.reshape(-1, x1.size(-1)).mean(0)[(None,) * (x1.dim() - 1)]
        x1_ = (x1 - mean).div(self.lengthscale)
        x2_ = (x2 - mean).div(self.lengthscale)
        distance = self.covar_dist(x1_, x2_, diag=diag, **params)
        exp_component = torch.exp(-math.sqrt(self.nu * 2) * distance)
        if self.nu == 0.5:
            constant_component = 1
        elif self.nu == 1.5:
            constant_component = (math.sqrt(3) * distance).add(1)
        else:
            constant_component = (math.sqrt(5) * distance).add(1).add(5.0 / 3.0 * distance ** 2)
        return constant_component * exp_component


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", Parameter(torch.tensor(0.0)))
        self.register_constraint("raw_outputscale", Positive())

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = torch.as_tensor(value).to(self.raw_outputscale)
        self.initialize(raw_outputscale=self.raw_outputscale_constraint.inverse_transform(value))

    def forward(self, x1, x2, diag=False, **params):
        orig_output = self.base_kernel.forward(x1, x2, diag=diag, **params)
        outputscales = self.outputscale
        if diag:
            return _lazy.delazify(orig_output) * outputscales
        if isinstance(orig_output, _lazy.LazyTensor):
            return orig_output.mul(outputscales)
        return orig_output * outputscales


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = ModuleList(kernels)

    def forward(self, x1, x2, diag=False, **params):
        res = None
        for kern in self.kernels:
            term = kern(x1, x2, diag=diag, **params)
            term = _lazy.delazify(term)
            res = term if res is None else res + term
        return res


class RFFKernel(Kernel):
    has_lengthscale = True

    def __init__(self, num_samples, num_dims=None, **kwargs):
        super().__init__(**kwargs)
        self.num_samples = num_samples
        if num_dims is not None:
            self._init_weights(num_dims, num_samples)

    def _init_weights(self, num_dims=None, num_samples=None, randn_weights=None):
        if num_dims is not None and num_samples is not None:
            d, D = num_dims, num_samples
        if randn_weights is None:
            randn_shape = torch.Size([*self._batch_shape, d, D])
            randn_weights = torch.randn(randn_shape, dtype=self.raw_lengthscale.dtype,
                                        device=self.raw_lengthscale.device)
        self.register_buffer("randn_weights", randn_weights)

    def _featurize(self, x, normalize=False):
        x = x.matmul(self.randn_weights / self.lengthscale.transpose(-1, -2))
        z = torch.cat([torch.cos(x), torch.sin(x)], dim=-1)
        if normalize:
            z = z / math.sqrt(self.num_samples)
        return z

    def forward(self, x1, x2, diag=False, last_dim_is_batch=False, **kwargs):
        num_dims = x1.size(-1)
        if not hasattr(self, "randn_weights"):
            self._init_weights(num_dims, self.num_samples)
        x1_eq_x2 = torch.equal(x1, x2)
        z1 = self._featurize(x1, normalize=False)
        z2 = z1 if x1_eq_x2 else self._featurize(x2, normalize=False)
        D = float(self.num_samples)
        if diag:
            return (z1 * z2).sum(-1) / D
        if x1_eq_x2:
            return _lazy.RootLazyTensor(z1 / math.sqrt(D))
        return _lazy.MatmulLazyTensor(z1 / D, z2.transpose(-1, -2))
