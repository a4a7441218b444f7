"""gpytorch.lazy restated (test infrastructure): every "lazy" tensor is an
eager wrapper round a dense 2-D torch tensor.  Method semantics follow
gpytorch==1.4.0's *Cholesky branch* (SURVEY.md Appendix A rows `add_jitter`,
`root_inv_decomposition`, `inv_matmul / inv_quad / inv_quad_logdet`).

Reference call sites served: src/models/cmp.py:59-67,98-103;
src/kernels/cme_aggregate_kernel.py:63-70; src/means/cme_aggregate_mean.py:52-55;
src/likelihoods/cmp_likelihood.py:55-78,95-116;
src/variational/variational_strategy.py:21-65;
src/models/prediction_strategies/cmp_prediction_strategy.py:43,56-60,154-178.
"""
import torch
from .utils.cholesky import psd_safe_cholesky


def _dense(x):
    return x.evaluate() if isinstance(x, LazyTensor) else x


class LazyTensor:
    """Dense-backed node.  Built either from a tensor or from a thunk that is
    evaluated (once, cached) on first use - this keeps gpytorch's deferred
    evaluation order, which matters for autograd: a kernel matrix created under
    `torch.no_grad()` (src/models/cmp.py:31) but first evaluated later still
    carries gradient (SURVEY.md §3.1 subtlety (ii))."""

    def __init__(self, tensor):
        if callable(tensor) and not torch.is_tensor(tensor) and not isinstance(tensor, LazyTensor):
            self._thunk, self._value = tensor, None
        else:
            self._thunk, self._value = None, _dense(tensor)

    @property
    def tensor(self):
        if self._value is None:
            self._value = _dense(self._thunk())
            self._thunk = None
        return self._value

    # -- basic protocol ------------------------------------------------
    def evaluate(self):
        return self.tensor

    def evaluate_kernel(self):
        return self

    @property
    def shape(self):
        return self.tensor.shape

    @property
    def dtype(self):
        return self.tensor.dtype

    @property
    def device(self):
        return self.tensor.device

    @property
    def requires_grad(self):
        return self.tensor.requires_grad

    def size(self, *a):
        return self.tensor.size(*a)

    def dim(self):
        return self.tensor.dim()

    ndimension = dim

    def numel(self):
        return self.tensor.numel()

    def __getitem__(self, idx):
        return LazyTensor(lambda: self.tensor[idx])

    def t(self):
        return LazyTensor(lambda: self.tensor.transpose(-1, -2))

    def transpose(self, a, b):
        return LazyTensor(lambda: self.tensor.transpose(a, b))

    def detach(self):
        return LazyTensor(self.tensor.detach())

    def clone(self):
        return LazyTensor(self.tensor.clone())

    def to(self, *a, **k):
        return LazyTensor(self.tensor.to(*a, **k))

    def cpu(self, *a, **k):
        return LazyTensor(self.tensor.cpu())

    def cuda(self, *a, **k):
        return LazyTensor(self.tensor.cuda(*a, **k))

    def double(self):
        return LazyTensor(self.tensor.double())

    def float(self):
        return LazyTensor(self.tensor.float())

    def diag(self):
        return self.tensor.diagonal(dim1=-2, dim2=-1)

    def sum(self, *a, **k):
        return self.tensor.sum(*a, **k)

    # -- algebra ---------------------------------------------------------
    def matmul(self, other):
        if isinstance(other, LazyTensor):
            return LazyTensor(lambda: self.tensor @ other.tensor)
        return self.tensor @ other

    __matmul__ = matmul

    def __rmatmul__(self, other):
        return other @ self.tensor

    def mul(self, other):
        return LazyTensor(lambda: self.tensor * _dense(other))

    __mul__ = mul
    __rmul__ = mul

    def __neg__(self):
        return LazyTensor(lambda: -self.tensor)

    def __add__(self, other):
        return LazyTensor(lambda: self.tensor + _dense(other))

    __radd__ = __add__

    def __sub__(self, other):
        return LazyTensor(lambda: self.tensor - _dense(other))

    def add_diag(self, diag):
        def thunk():
            d = diag
            if d.dim() == 0 or d.numel() == 1:
                d = d.reshape(()).expand(self.tensor.size(-1))
            return self.tensor + torch.diag_embed(d)
        return LazyTensor(thunk)

    def add_jitter(self, jitter_val=1e-3):
        def thunk():
            n = self.tensor.size(-1)
            eye = torch.eye(n, dtype=self.tensor.dtype, device=self.tensor.device)
            return self.tensor + jitter_val * eye
        return LazyTensor(thunk)

    # -- decompositions (Cholesky branch) ---------------------------------
    def cholesky(self, upper=False):
        L = psd_safe_cholesky(self.tensor)
        return LazyTensor(L.transpose(-1, -2) if upper else L)

    def root_decomposition(self):
        return RootLazyTensor(psd_safe_cholesky(self.tensor))

    def root_inv_decomposition(self, initial_vectors=None, test_vectors=None):
        # root = (L^{-1})^T, so that root @ root.T = A^{-1}
        L = psd_safe_cholesky(self.tensor)
        eye = torch.eye(L.size(-1), dtype=L.dtype, device=L.device)
        Linv = torch.linalg.solve_triangular(L, eye, upper=False)
        return RootLazyTensor(Linv.transpose(-1, -2))

    def inv_matmul(self, right_tensor, left_tensor=None):
        squeeze = right_tensor.dim() == 1
        rhs = right_tensor.unsqueeze(-1) if squeeze else right_tensor
        L = psd_safe_cholesky(self.tensor)
        sol = torch.cholesky_solve(rhs, L)
        if left_tensor is not None:
            sol = left_tensor @ sol
        return sol.squeeze(-1) if squeeze else sol

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        L = psd_safe_cholesky(self.tensor)
        inv_quad_term = None
        if inv_quad_rhs is not None:
            rhs = inv_quad_rhs.unsqueeze(-1) if inv_quad_rhs.dim() == 1 else inv_quad_rhs
            half = torch.linalg.solve_triangular(L, rhs, upper=False)
            inv_quad_term = (half * half).sum(-2)
            if reduce_inv_quad:
                inv_quad_term = inv_quad_term.sum(-1)
        logdet_term = None
        if logdet:
            logdet_term = 2.0 * L.diagonal(dim1=-2, dim2=-1).log().sum(-1)
        return inv_quad_term, logdet_term

    def inv_quad(self, tensor, reduce_inv_quad=True):
        return self.inv_quad_logdet(tensor, logdet=False, reduce_inv_quad=reduce_inv_quad)[0]

    def logdet(self):
        return self.inv_quad_logdet(None, logdet=True)[1]


class NonLazyTensor(LazyTensor):
    pass


class DiagLazyTensor(LazyTensor):
    def __init__(self, diag):
        self._diag = diag
        super().__init__(torch.diag_embed(diag))

    def diag(self):
        return self._diag


class RootLazyTensor(LazyTensor):
    def __init__(self, root):
        root = _dense(root)
        self._root = root
        super().__init__(root @ root.transpose(-1, -2))

    @property
    def root(self):
        return LazyTensor(self._root)


class LowRankRootLazyTensor(RootLazyTensor):
    pass


class MatmulLazyTensor(LazyTensor):
    def __init__(self, left, right):
        super().__init__(lambda: _dense(left) @ _dense(right))


class SumLazyTensor(LazyTensor):
    def __init__(self, *tensors):
        def thunk():
            out = _dense(tensors[0])
            for t in tensors[1:]:
                out = out + _dense(t)
            return out
        super().__init__(thunk)


class LazyEvaluatedKernelTensor(LazyTensor):
    """k(x1, x2) evaluated on first use (gpytorch `lazily_evaluate_kernels`)."""

    def __init__(self, x1, x2, kernel, **params):
        self.x1, self.x2, self.kernel, self.params = x1, x2, kernel, params
        super().__init__(lambda: _dense(kernel.forward(x1, x2, **params)))


class TriangularLazyTensor(LazyTensor):
    def __init__(self, tensor, upper=False):
        super().__init__(tensor)
        self.upper = upper

    def inv_matmul(self, right_tensor, left_tensor=None):
        squeeze = right_tensor.dim() == 1
        rhs = right_tensor.unsqueeze(-1) if squeeze else right_tensor
        sol = torch.linalg.solve_triangular(self.tensor, rhs, upper=self.upper)
        if left_tensor is not None:
            sol = left_tensor @ sol
        return sol.squeeze(-1) if squeeze else sol


class CholLazyTensor(RootLazyTensor):
    pass


def lazify(obj):
    if isinstance(obj, LazyTensor):
        return obj
    return NonLazyTensor(obj)


def delazify(obj):
    return _dense(obj)
