"""Minimal matrix-like wrappers returned where the reference hands out gpytorch lazy
tensors (`.evaluate()`, `.diag()`, `.matmul()`, `.cpu()`, scalar `*`): only what the
unchanged trainers touch (SURVEY.md §8b "Returned distribution object ...")."""
import torch

from . import functional as F


class DenseMatrix:
    def __init__(self, tensor):
        self.tensor = tensor

    @property
    def shape(self):
        return self.tensor.shape

    def size(self, i=None):
        return self.tensor.size() if i is None else self.tensor.size(i)

    def evaluate(self):
        return self.tensor

    def diag(self):
        return self.tensor.diagonal()

    def matmul(self, W):
        return F.matmul(self.tensor, W)

    __matmul__ = matmul

    def t(self):
        return DenseMatrix(self.tensor.t())

    def add_diag(self, diag):
        return DenseMatrix(self.tensor + torch.diag_embed(diag.expand(self.tensor.size(-1))))

    def add_jitter(self, jitter_val=1e-3):
        n = self.tensor.size(-1)
        return DenseMatrix(self.tensor + jitter_val * torch.eye(n, dtype=self.tensor.dtype, device=self.tensor.device))

    def __add__(self, other):
        return DenseMatrix(self.tensor + (other.evaluate() if hasattr(other, "evaluate") else other))

    def __mul__(self, c):
        return DenseMatrix(self.tensor * c)

    __rmul__ = __mul__

    def detach(self):
        return DenseMatrix(self.tensor.detach())

    def cpu(self):
        return DenseMatrix(self.tensor.cpu())

    def to(self, *a, **k):
        return DenseMatrix(self.tensor.to(*a, **k))


class DiagLazyTensor(DenseMatrix):
    """Diagonal matrix given by its diagonal (gpytorch.lazy.DiagLazyTensor as used by src/kernels/delta_kernel.py:11
    and src/models/bagged_gp.py:78)."""

    def __init__(self, diag):
        self._diag = diag

    @property
    def tensor(self):
        return torch.diag_embed(self._diag)

    @property
    def shape(self):
        return torch.Size([self._diag.numel(), self._diag.numel()])

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def diag(self):
        return self._diag

    def matmul(self, W):
        return self._diag.unsqueeze(-1) * W if W.dim() == 2 else self._diag * W

    __matmul__ = matmul


class SumLazyTensor(DenseMatrix):
    """Sum of matrix-likes, evaluated on demand (src/variational/variational_strategy.py:64-66)."""

    def __init__(self, *terms):
        self.terms = [lazify(t) for t in terms]

    @property
    def tensor(self):
        out = self.terms[0].evaluate()
        for t in self.terms[1:]:
            out = out + t.evaluate()
        return out

    @property
    def shape(self):
        return self.terms[0].shape

    def diag(self):
        out = self.terms[0].diag()
        for t in self.terms[1:]:
            out = out + t.diag()
        return out


class MatmulLazyTensor(DenseMatrix):
    """Product of two matrix-likes, evaluated on demand on the tensor cores."""

    def __init__(self, left, right):
        self.left, self.right = lazify(left), lazify(right)

    @property
    def tensor(self):
        return F.matmul(self.left.evaluate(), self.right.evaluate())

    @property
    def shape(self):
        return torch.Size([self.left.shape[0], self.right.shape[1]])

    def diag(self):
        return (self.left.evaluate() * self.right.evaluate().t()).sum(-1)


class RootLazyTensor(DenseMatrix):
    """R R^T given its root (low-rank roots of the RFF kernels, src/kernels/rff_kernel.py:104-110)."""

    def __init__(self, root):
        self.root = delazify(root)

    @property
    def tensor(self):
        return F.matmul(self.root, self.root, transb=True)

    @property
    def shape(self):
        return torch.Size([self.root.size(0), self.root.size(0)])

    def diag(self):
        return (self.root * self.root).sum(-1)


class LazyCovariance:
    """Covariance known through callables: dense() and diag() (posterior covariances that are
    only materialised when somebody asks, e.g. the 2000-point NLL chunks of
    experiments/swiss_roll/core/metrics.py:19-32)."""

    def __init__(self, dense_fn, diag_fn, n, scale=1.0):
        self._dense_fn, self._diag_fn, self.n, self.scale = dense_fn, diag_fn, n, scale

    @property
    def shape(self):
        return torch.Size([self.n, self.n])

    def evaluate(self):
        return self._dense_fn() * self.scale

    def diag(self):
        return self._diag_fn() * self.scale

    def __mul__(self, c):
        return LazyCovariance(self._dense_fn, self._diag_fn, self.n, self.scale * c)

    __rmul__ = __mul__

    def cpu(self):
        dense_fn, diag_fn = self._dense_fn, self._diag_fn
        return LazyCovariance(lambda: dense_fn().cpu(), lambda: diag_fn().cpu(), self.n, self.scale)

    def detach(self):
        return self

    def add_jitter(self, jitter_val=1e-3):
        dense_fn, diag_fn, n = self._dense_fn, self._diag_fn, self.n

        def dense():
            K = dense_fn()
            return K + jitter_val * torch.eye(n, dtype=K.dtype, device=K.device)
        return LazyCovariance(dense, lambda: diag_fn() + jitter_val, n, self.scale)


def lazify(obj):
    if hasattr(obj, "evaluate"):
        return obj
    return DenseMatrix(obj)


def delazify(obj):
    return obj.evaluate() if hasattr(obj, "evaluate") else obj
