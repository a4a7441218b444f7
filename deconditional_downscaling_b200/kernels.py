"""Kernel modules with the reference's / gpytorch's class API, evaluated by libdcgp.

Mirrors `src/kernels/__init__.py:1-5` (CMEAggregateKernel, DeltaKernel, RFFKernel) and the
gpytorch kernels the reference's model builders instantiate
(experiments/swiss_roll/models/exact_cmp.py:32-40,
experiments/downscaling/models/variational_cmp.py:36-55): RBFKernel, MaternKernel,
ScaleKernel, `k1 + k2` -> AdditiveKernel with `.kernels`, gpytorch's RBF RFFKernel.

Parameter conventions are gpytorch's: raw (pre-softplus)
nn.Parameters `raw_lengthscale` (1, ard_num_dims) and `raw_outputscale` (0-dim) with
`lengthscale` / `outputscale` properties and `initialize(raw_x=...)`.

A kernel tree is flattened into additive terms; all stationary terms are evaluated by ONE
launch of the fused Gram kernel (ops.gram), RFF terms by feature maps + tensor-core GEMMs.
"""
import math
from typing import List, Optional

import numpy as np
import torch
from torch import nn
from torch.nn.functional import softplus

from . import functional as F
from .functional import KernelMeta
from .constraints import Positive
from .module import Module


def _as_2d(x):
    return x.unsqueeze(-1) if x.dim() == 1 else x


# ------------------------------------------------------------------------------------------
# additive terms produced by flattening a kernel tree
# ------------------------------------------------------------------------------------------
class _Term:
    def __init__(self, kind, dims, lengthscale, outputscale, weights=None):
        self.kind, self.dims, self.lengthscale, self.outputscale, self.weights = kind, dims, lengthscale, outputscale, weights


class KernelOps:
    """Evaluation plan of a kernel module for inputs of width `d` and dtype `dtype`."""

    def __init__(self, kernel: "Kernel", d: int, dtype, device):
        terms = kernel._terms(d, None)
        self.dtype, self.device = dtype, device
        self.flat = [t for t in terms if t.kind in ("rbf", "matern32", "matern12", "matern52")]
        self.rff = [t for t in terms if t.kind == "rff"]
        self.delta = [t for t in terms if t.kind == "delta"]
        if len(self.flat) > 4:
            raise NotImplementedError("at most 4 stationary additive terms are fused into one Gram launch")
        self.meta = KernelMeta(tuple(t.kind for t in self.flat), tuple(tuple(t.dims) for t in self.flat))
        self.hyp = []
        for t in self.flat:
            self.hyp += [t.lengthscale.reshape(-1).to(dtype), t.outputscale.reshape(1).to(dtype)]

    # -- helpers -----------------------------------------------------------------------------
    def _meta(self, diag_const=0.0):
        return KernelMeta(self.meta.kinds, self.meta.dims, float(diag_const))

    def _delta_scale(self):
        s = None
        for t in self.delta:
            v = t.outputscale.reshape(1).to(self.dtype)
            s = v if s is None else s + v
        return s

    def _features(self, t, x):
        from .rff import rff_features
        return rff_features(x[:, list(t.dims)].contiguous(), t.weights.to(self.dtype), t.lengthscale.reshape(-1).to(self.dtype))

    @property
    def is_flat(self):
        return not self.rff and not self.delta

    # -- dense Gram ---------------------------------------------------------------------------
    def gram(self, x1, x2=None, diag_const=0.0):
        """Dense k(x1, x2); x2 None means k(x1, x1) (+ diag_const, + delta terms on the diagonal)."""
        x1 = _as_2d(x1)
        x2 = None if x2 is None else _as_2d(x2)
        out = None
        dd = self._delta_scale() if x2 is None else None
        if self.flat:
            out = F.gram(self._meta(diag_const if x2 is None else 0.0), self.hyp, x1, x2, diag_dev=dd)
        for t in self.rff:
            D = t.weights.shape[1]
            p1 = self._features(t, x1)
            p2 = p1 if x2 is None else self._features(t, x2)
            k = F.matmul(p1, p2, transb=True) * (t.outputscale.to(self.dtype) / D)
            out = k if out is None else out + k
        if not self.flat and x2 is None and (dd is not None or diag_const):
            eye = torch.eye(x1.size(0), dtype=self.dtype, device=x1.device)
            out = out + eye * ((dd if dd is not None else 0.0) + diag_const)
        if out is None:  # pure delta kernel
            if x2 is None:
                return torch.eye(x1.size(0), dtype=self.dtype, device=x1.device) * (dd + diag_const)
            return torch.zeros(x1.size(0), x2.size(0), dtype=self.dtype, device=x1.device)
        return out

    def diag(self, x):
        x = _as_2d(x)
        out = torch.zeros(x.size(0), dtype=self.dtype, device=x.device)
        for t in self.flat:
            out = out + t.outputscale.reshape(()).to(self.dtype)
        for t in self.rff:
            p = self._features(t, x)
            out = out + (p * p).sum(-1) * (t.outputscale.to(self.dtype) / t.weights.shape[1])
        for t in self.delta:
            out = out + t.outputscale.reshape(()).to(self.dtype)
        return out

    # -- products without keeping K ------------------------------------------------------------
    def matmul(self, x1, x2, W, diag_const=0.0):
        """(k(x1, x2) [+ diag terms when x2 is None]) @ W."""
        x1 = _as_2d(x1)
        x2 = None if x2 is None else _as_2d(x2)
        Wm = _as_2d(W)
        out = None
        dd = self._delta_scale() if x2 is None else None
        if self.flat:
            out = F.kernel_matmul(self._meta(diag_const if x2 is None else 0.0), self.hyp, x1, x2, Wm, diag_dev=dd)
        elif x2 is None and (dd is not None or diag_const):
            out = Wm * ((dd if dd is not None else 0.0) + diag_const)
        for t in self.rff:
            D = t.weights.shape[1]
            p1 = self._features(t, x1)
            p2 = p1 if x2 is None else self._features(t, x2)
            k = F.matmul(p1, F.matmul(p2, Wm, transa=True)) * (t.outputscale.to(self.dtype) / D)
            out = k if out is None else out + k
        return out.squeeze(-1) if W.dim() == 1 else out

    # -- bag aggregation ---------------------------------------------------------------------------
    def bagsum(self, z, x, offsets):
        """out[a, i] = sum_{j in bag a} k(z_i, x_j)."""
        z, x = _as_2d(z), _as_2d(x)
        out = None
        if self.flat:
            out = F.gram_bagsum(self.meta, self.hyp, z, x, offsets)
        for t in self.rff:
            D = t.weights.shape[1]
            pb = F.bag_sum(self._features(t, x), offsets)                     # n_bags x 2D
            k = F.matmul(pb, self._features(t, z), transb=True) * (t.outputscale.to(self.dtype) / D)
            out = k if out is None else out + k
        return out

    def bag_blocksum(self, x, offsets):
        """out[a] = sum_{i,j in bag a} k(x_i, x_j) (delta terms add n_a * scale)."""
        x = _as_2d(x)
        sizes = (offsets[1:] - offsets[:-1]).to(self.dtype)
        out = torch.zeros(offsets.numel() - 1, dtype=self.dtype, device=x.device)
        if self.flat:
            out = out + F.gram_bag_blocksum(self.meta, self.hyp, x, offsets)
        for t in self.rff:
            pb = F.bag_sum(self._features(t, x), offsets)
            out = out + (pb * pb).sum(-1) * (t.outputscale.to(self.dtype) / t.weights.shape[1])
        dd = self._delta_scale()
        if dd is not None:
            out = out + dd * sizes
        return out

    # -- (k(x,x) + c I)^{-1} B ----------------------------------------------------------------------
    def factor(self, x, diag_const):
        """Returns a solver for (k(x,x) + diag_const I); flat kernels use the cached-Cholesky +
        low-rank-adjoint path, others the differentiable dense Cholesky."""
        x = _as_2d(x)
        return KernelInverse(self, x, float(diag_const))


class KernelInverse:
    """Represents Lambda^{-1} for Lambda = k(x,x) + c I; plays the role of the reference's
    `root_inv_*_covar` (R R^T = Lambda^{-1}, src/models/cmp.py:66-67) without ever forming R."""

    def __init__(self, kops: KernelOps, x, diag_const, detach=False):
        self.kops, self.x, self.diag_const = kops, x, diag_const
        self.n = x.size(0)
        if kops.is_flat:
            self.meta = kops._meta(diag_const)
            self.hyp = [h.detach() for h in kops.hyp] if detach else kops.hyp
            self.L = F.kernel_factor(self.meta, self.hyp, x)
        else:
            K = kops.gram(x, None, diag_const=diag_const)
            if detach:
                K = K.detach()
            self.L = F.psd_safe_cholesky(K)
        self._cache = {}

    def solve(self, B):
        """Lambda^{-1} B, differentiable w.r.t. B and the kernel hyper-parameters."""
        if self.kops.is_flat:
            return F.kernel_solve(self.meta, self.hyp, self.x, self.L, B)
        return F.chol_solve(self.L, B)

    def cached(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    def evaluate_root(self):
        """Explicit R = L^{-T} (only for callers that insist on the reference's root form)."""
        eye = torch.eye(self.n, dtype=self.L.dtype, device=self.L.device)
        return F.trsm(self.L, eye, trans=False).t()

    # moved by hand in ExactCMP.to()/cpu() in the reference (src/models/exact_cmp.py:157-169)
    def to(self, *a, **k):
        return self

    def cpu(self, *a, **k):
        return self


class KernelMatrix:
    """k(x1, x2) kept symbolic (the analogue of gpytorch's lazily evaluated kernel tensor):
    dense on `.evaluate()`, never stored by `.matmul(W)`."""

    def __init__(self, kernel, x1, x2=None, diag_const=0.0, extra=None):
        self.kernel, self.x1, self.x2, self.diag_const = kernel, _as_2d(x1), None if x2 is None else _as_2d(x2), diag_const
        self.extra = list(extra or [])   # further KernelMatrix terms over the same inputs (k + noise_kernel)

    def _ops(self):
        return KernelOps(self.kernel, self.x1.size(-1), self.x1.dtype, self.x1.device)

    @property
    def shape(self):
        return torch.Size([self.x1.size(0), (self.x1 if self.x2 is None else self.x2).size(0)])

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def evaluate(self):
        out = self._ops().gram(self.x1, self.x2, diag_const=self.diag_const)
        for e in self.extra:
            out = out + e.evaluate()
        return out

    def matmul(self, W):
        out = self._ops().matmul(self.x1, self.x2, W, diag_const=self.diag_const)
        for e in self.extra:
            out = out + e.matmul(W)
        return out

    __matmul__ = matmul

    def diag(self):
        out = self._ops().diag(self.x1) + self.diag_const
        for e in self.extra:
            out = out + e.diag()
        return out

    def add_diag(self, diag):
        c = float(diag.reshape(-1)[0]) if torch.is_tensor(diag) else float(diag)
        return KernelMatrix(self.kernel, self.x1, self.x2, self.diag_const + c, self.extra)

    def add_jitter(self, jitter_val=1e-3):
        return KernelMatrix(self.kernel, self.x1, self.x2, self.diag_const + jitter_val, self.extra)

    def __add__(self, other):
        if isinstance(other, KernelMatrix):
            return KernelMatrix(self.kernel, self.x1, self.x2, self.diag_const, self.extra + [other])
        return self.evaluate() + (other.evaluate() if hasattr(other, "evaluate") else other)

    def t(self):
        if self.x2 is None:
            return self
        return KernelMatrix(self.kernel, self.x2, self.x1, self.diag_const, [e.t() for e in self.extra])

    def root_inv_decomposition(self):
        if self.x2 is not None or self.extra:
            raise NotImplementedError
        return _RootHolder(self._ops().factor(self.x1, self.diag_const))

    def to(self, *a, **k):
        return self

    def cpu(self, *a, **k):
        return self

    def detach(self):
        return self


class _RootHolder:
    """`foo.root_inv_decomposition().root` spelling of the reference (src/models/cmp.py:67)."""

    def __init__(self, inv):
        self.root = inv


# ------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------
class Kernel(Module):
    has_lengthscale = False
    kind = None

    def __init__(self, ard_num_dims=None, active_dims=None, batch_shape=None, lengthscale_prior=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor(list(active_dims), dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        # host copy: the column selection shapes every launch and must not cost a device read-back each time
        self._active_dims_host = None if active_dims is None else [int(i) for i in active_dims.tolist()]
        self.ard_num_dims = ard_num_dims
        if self.has_lengthscale:
            nd = 1 if ard_num_dims is None else ard_num_dims
            self.register_parameter("raw_lengthscale", nn.Parameter(torch.zeros(1, nd)))
            self.raw_lengthscale_constraint = Positive()

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale) if self.has_lengthscale else None

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value).to(self.raw_lengthscale)
        self.initialize(raw_lengthscale=value + torch.log(-torch.expm1(-value)))

    def _dims(self, d, parent_dims, own=True):
        if own and self.active_dims is not None:
            if self._active_dims_host is None or len(self._active_dims_host) != self.active_dims.numel():
                self._active_dims_host = [int(i) for i in self.active_dims.tolist()]   # e.g. after load_state_dict
            own = self._active_dims_host
            return own if parent_dims is None else [parent_dims[i] for i in own]
        return list(range(d)) if parent_dims is None else list(parent_dims)

    def _terms(self, d, parent_dims, scale=None, own=True) -> List[_Term]:
        """Flatten into additive terms over the ORIGINAL input columns (`own`: apply this
        kernel's own active_dims; False when a ScaleKernel already applied them)."""
        if self.kind is None:
            raise NotImplementedError(f"{type(self).__name__} cannot be flattened")
        dims = self._dims(d, parent_dims, own)
        ls = self.lengthscale.reshape(-1)
        if ls.numel() == 1 and len(dims) > 1:
            ls = ls.expand(len(dims))
        os_ = scale if scale is not None else torch.ones(1, dtype=ls.dtype, device=ls.device)
        return [_Term(self.kind, dims, ls, os_)]

    def ops(self, x):
        x = _as_2d(x)
        return KernelOps(self, x.size(-1), x.dtype, x.device)

    def forward(self, x1, x2=None, diag=False, **params):
        x1 = _as_2d(x1)
        same = x2 is None or (x2 is x1) or (x1.shape == _as_2d(x2).shape and torch.equal(x1, _as_2d(x2)))
        if diag:
            return self.ops(x1).diag(x1)
        return self.ops(x1).gram(x1, None if same else x2)

    def __call__(self, x1, x2=None, diag=False, **params):
        x1 = _as_2d(x1)
        if diag:
            return self.ops(x1).diag(x1)
        if x2 is not None:
            x2 = _as_2d(x2)
            if x2 is x1 or (x1.shape == x2.shape and x1.data_ptr() == x2.data_ptr()):
                x2 = None
        return KernelMatrix(self, x1, x2)

    def __add__(self, other):
        ks = (list(self.kernels) if isinstance(self, AdditiveKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, AdditiveKernel) else [other])
        return AdditiveKernel(*ks)


class RBFKernel(Kernel):
    has_lengthscale = True
    kind = "rbf"


class MaternKernel(Kernel):
    has_lengthscale = True

    def __init__(self, nu=2.5, **kwargs):
        if nu not in (0.5, 1.5, 2.5):
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super().__init__(**kwargs)
        self.nu = nu
        self.kind = {0.5: "matern12", 1.5: "matern32", 2.5: "matern52"}[nu]


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        if base_kernel.active_dims is not None:
            kwargs["active_dims"] = base_kernel.active_dims
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", nn.Parameter(torch.tensor(0.0)))
        self.raw_outputscale_constraint = Positive()

    @property
    def outputscale(self):
        return softplus(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = torch.as_tensor(value).to(self.raw_outputscale)
        self.initialize(raw_outputscale=value + torch.log(-torch.expm1(-value)))

    def _terms(self, d, parent_dims, scale=None, own=True):
        s = self.outputscale.reshape(1)
        s = s if scale is None else s * scale
        # like gpytorch, ScaleKernel inherits the base kernel's active_dims and calls its
        # forward directly, so the base kernel's own active_dims are not applied twice
        return self.base_kernel._terms(d, self._dims(d, parent_dims, own), s, own=False)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)

    def _terms(self, d, parent_dims, scale=None, own=True):
        out = []
        dims = None if (parent_dims is None and self.active_dims is None) else self._dims(d, parent_dims, own)
        for k in self.kernels:
            out += k._terms(d, dims, scale)
        return out


class DeltaKernel(Kernel):
    """k(x1, x2) = delta(x1, x2) (src/kernels/delta_kernel.py:5-14).  Only its diagonal form is
    needed on the hot path: ScaleKernel(DeltaKernel) = sigma_ind^2 I added to k(x, x)
    (src/models/cmp.py:61-62), and its outputscale in CMPLikelihood (variational_cmp.py:110-111)."""
    kind = "delta"

    def _terms(self, d, parent_dims, scale=None, own=True):
        os_ = scale if scale is not None else torch.ones(1)
        return [_Term("delta", [], None, os_)]

    def forward(self, x1, x2=None, **params):
        x1 = _as_2d(x1)
        if x2 is None or torch.equal(x1, _as_2d(x2)):
            return torch.eye(x1.size(0), dtype=x1.dtype, device=x1.device)
        dist = torch.cdist(x1, _as_2d(x2))
        return (dist < torch.finfo(torch.float16).eps).to(x1.dtype)


class StandardMultivariateTStudent(torch.distributions.Distribution):
    """Zero-mean identity-scale multivariate Student-t sampler (src/kernels/rff_kernel.py:7-29):
    x * rsqrt(chi2(df) / df) with one chi-square draw per sample row."""
    arg_constraints = {}
    has_rsample = True

    def __init__(self, df, validate_args=None):
        self.df = torch.as_tensor(float(df))
        self._chi2 = torch.distributions.Chi2(self.df)
        super().__init__(self.df.size(), validate_args=False)

    def rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(sample_shape)
        x = torch.randn(*shape).view(shape[0], -1)
        z = self._chi2.rsample((shape[0],)).unsqueeze(-1)
        return (x * torch.rsqrt(z / self.df)).view(*shape)


class RFFKernel(Kernel):
    """Random Fourier feature kernel, RBF (nu=inf, gpytorch's RFFKernel) or Matern (Student-t
    spectral sampling) — src/kernels/rff_kernel.py:32-114.  Frequencies live in the buffer
    `rand_weights` (d, D) (`randn_weights` for the plain gpytorch kernel), drawn at first use and
    never trained.  Phi = [cos(x W / l^T), sin(.)]; k(x1, x2) = Phi1 Phi2^T / D."""
    has_lengthscale = True
    kind = "rff"
    _buffer_name = "rand_weights"

    def __init__(self, num_samples, num_dims=None, nu=np.inf, **kwargs):
        super().__init__(**kwargs)
        self.num_samples, self.nu = num_samples, nu
        if num_dims is not None:
            self._init_weights(num_dims, num_samples)

    def _sample(self, d, D):
        if self.nu == np.inf:
            return torch.randn(d, D)
        # reference call order: sample_shape (D, d) then transpose (rff_kernel.py:65-80)
        return StandardMultivariateTStudent(df=2 * self.nu).sample(torch.Size([D, d])).t()

    def _init_weights(self, num_dims=None, num_samples=None, rand_weights=None):
        if rand_weights is None:
            rand_weights = self._sample(num_dims, num_samples)
        rand_weights = rand_weights.to(dtype=self.raw_lengthscale.dtype, device=self.raw_lengthscale.device)
        self.register_buffer(self._buffer_name, rand_weights.contiguous())

    @property
    def weights(self):
        return getattr(self, self._buffer_name, None)

    def _terms(self, d, parent_dims, scale=None, own=True):
        dims = self._dims(d, parent_dims, own)
        if self.weights is None:
            self._init_weights(len(dims), self.num_samples)
        ls = self.lengthscale.reshape(-1)
        if ls.numel() == 1 and len(dims) > 1:
            ls = ls.expand(len(dims))
        os_ = scale if scale is not None else torch.ones(1, dtype=ls.dtype, device=ls.device)
        return [_Term("rff", dims, ls, os_, self.weights)]


class GPyTorchRFFKernel(RFFKernel):
    """gpytorch.kernels.RFFKernel (RBF spectral density, buffer `randn_weights`) as instantiated at
    experiments/downscaling/models/variational_cmp.py:39."""
    _buffer_name = "randn_weights"

    def __init__(self, num_samples, num_dims=None, **kwargs):
        super().__init__(num_samples=num_samples, num_dims=num_dims, nu=np.inf, **kwargs)


def cme_weights(inv: "KernelInverse", bag_kernel, bags_values, x):
    """W = Lambda^{-1} l(bags_values, x)  (N x n), the CME weights both CMEAggregateMean and
    CMEAggregateKernel need; computed once per (x, refresh) through the inverse's cache."""
    x = _as_2d(x)
    key = ("W", x.data_ptr(), tuple(x.shape), x._version, torch.is_grad_enabled())
    bv = _as_2d(bags_values)
    return inv.cached(key, lambda: inv.solve(bag_kernel.ops(bv).gram(bv, x)))


class CMEAggregateKernel(Kernel):
    """Q = B1^T Lambda^{-1} K Lambda^{-1} B2 with B_i = l(bags_values, x_i)
    (src/kernels/cme_aggregate_kernel.py:4-70).  `individuals_covar` is a KernelMatrix
    (never materialised: K W goes through KernelMatrix.matmul) and `root_inv_bags_covar` a
    KernelInverse; Q = W1^T (K W2), W = Lambda^{-1} B, costs O(N^2 n) instead of the
    reference's O(N^2 k) products with an explicit root."""

    def __init__(self, bag_kernel, bags_values, individuals_covar, root_inv_bags_covar):
        super().__init__()
        self.bag_kernel = bag_kernel
        self.register_buffer("bags_values", bags_values)
        self.individuals_covar = individuals_covar
        self.root_inv_bags_covar = root_inv_bags_covar

    def weights(self, x):
        return cme_weights(self.root_inv_bags_covar, self.bag_kernel, self.bags_values, x)

    def forward(self, x1, x2=None, **kwargs):
        W1 = self.weights(x1)
        W2 = W1 if (x2 is None or x2 is x1) else self.weights(x2)
        KW2 = self.individuals_covar.matmul(W2)
        return F.matmul(W1, KW2, transa=True)

    def __call__(self, x1, x2=None, **kwargs):
        from .lazy import DenseMatrix
        return DenseMatrix(self.forward(x1, x2))
