"""Differentiable building blocks of the hot path: torch.autograd.Function
wrappers whose forward AND backward are libdcgp kernels (ops.py).

Backward passes are written by hand so that no O(N^3) adjoint is ever formed
for the big bag-covariance solve: `kernel_solve` returns W = (l(y~,y~)+cI)^{-1}B
and back-propagates the rank-m adjoint -(Lambda^{-1} gW) W^T straight into the
Gram-shaped hyper-parameter reduction.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import ops
from .ops import FlatSpec


class CachingError(RuntimeError):
    """gpytorch.utils.errors.CachingError."""


class NotPSDError(RuntimeError):
    """Raised when a Cholesky factorisation meets a non-positive pivot (same
    name and role as gpytorch.utils.errors.NotPSDError, which the reference
    harness catches at experiments/swiss_roll/core/metrics.py:30)."""


@dataclass(frozen=True)
class KernelMeta:
    kinds: Tuple[str, ...]
    dims: Tuple[Tuple[int, ...], ...]
    diag_const: float = 0.0


def _spec(meta: KernelMeta, hyp: Sequence[torch.Tensor]) -> FlatSpec:
    return FlatSpec(meta.kinds, meta.dims, [h.detach() for h in hyp[0::2]], [h.detach() for h in hyp[1::2]])


def _hyper_grads(meta: KernelMeta, hyp, dls, dos, needs):
    out = []
    for c in range(len(meta.kinds)):
        ls, os_ = hyp[2 * c], hyp[2 * c + 1]
        out.append(dls[c, :ls.numel()].reshape(ls.shape).to(ls.dtype) if needs[2 * c] else None)
        out.append(dos[c].reshape(os_.shape).to(os_.dtype) if needs[2 * c + 1] else None)
    return out


# ---------------------------------------------------------------------- Gram
class _GramFn(Function):
    @staticmethod
    def forward(ctx, meta, x1, x2, diag_dev, *hyp):
        spec = _spec(meta, hyp)
        K = ops.gram(spec, x1, x2, diag_dev=diag_dev, diag_const=meta.diag_const)
        ctx.meta, ctx.same = meta, x2 is None
        ctx.save_for_backward(x1, x2 if x2 is not None else x1, *hyp)
        return K

    @staticmethod
    def backward(ctx, G):
        x1, x2, *hyp = ctx.saved_tensors
        meta, same = ctx.meta, ctx.same
        spec = _spec(meta, hyp)
        G = G.contiguous()
        need = ctx.needs_input_grad
        dls, dos, dx1 = ops.gram_bwd(spec, x1, None if same else x2, G, need_dx1=need[1])
        dx2 = None
        if not same and need[2]:
            dx2 = ops.gram_bwd(spec, x2, x1, G.t().contiguous(), need_dx1=True)[2]
        ddiag = G.diagonal().sum().reshape(1) if (need[3] and same) else None
        return (None, dx1, dx2, ddiag, *_hyper_grads(meta, hyp, dls, dos, need[4:]))


def gram(meta: KernelMeta, hyp: Sequence[torch.Tensor], x1, x2=None, diag_dev=None):
    """Dense k(x1, x2) (+ diag on x2 is None).  hyp = [ls_0, os_0, ls_1, os_1, ...]."""
    return _GramFn.apply(meta, x1, x2, diag_dev, *hyp)


# ---------------------------------------------------------------------- GEMM
class _MatmulFn(Function):
    @staticmethod
    def forward(ctx, A, B, ta, tb, precise):
        ctx.ta, ctx.tb = ta, tb
        ctx.mm = ops.gemm_x3 if precise else ops.gemm
        ctx.save_for_backward(A, B)
        return ctx.mm(A, B, transa=ta, transb=tb)

    @staticmethod
    def backward(ctx, gC):
        A, B = ctx.saved_tensors
        ta, tb, mm = ctx.ta, ctx.tb, ctx.mm
        gC = gC.contiguous()
        gA = gB = None
        if ctx.needs_input_grad[0]:
            gA = mm(B, gC, transa=tb, transb=True) if ta else mm(gC, B, transa=False, transb=not tb)
        if ctx.needs_input_grad[1]:
            gB = mm(gC, A, transa=True, transb=ta) if tb else mm(A, gC, transa=not ta, transb=False)
        return gA, gB, None, None, None


def matmul(A, B, transa=False, transb=False, precise=False):
    """op(A) @ op(B) on the tensor cores; 1-D B is treated as a column.  `precise`: FP32 operands are multiplied
    with 3xTF32 (~FP32 accuracy) whatever the size - for products whose result enters a cancellation."""
    vec = B.dim() == 1
    if vec:
        B = B.unsqueeze(-1)
    out = _MatmulFn.apply(A, B, transa, transb, precise)
    return out.squeeze(-1) if vec else out


# ---------------------------------------------------------------------- Cholesky & solves
_DEFERRED = None   # list collecting the device `info` words of the factorisations of a deferred region
_POOL = None       # InfoPool handing out those words from one buffer (trainer steps), or None


class InfoPool:
    """One device buffer of int32 failure words: every factorisation of a deferred region writes its LAPACK-style
    `info` into the next slot, so that the whole step's status is ONE tensor - the veto argument of the device-side
    optimiser update, the payload of the cross-rank MAX all-reduce and a single host read."""

    def __init__(self, device, capacity=64):
        self.buf = torch.zeros(capacity, dtype=torch.int32, device=device)
        self.used = 0

    def reset(self):
        self.used = 0

    def next(self, words=1):
        if self.used + words > self.buf.numel():
            raise RuntimeError("InfoPool exhausted: more factorisations in one step than slots")
        slot = self.buf[self.used:self.used + words]
        self.used += words
        return slot

    def words(self):
        return self.buf[:self.used]

    def any_failed(self):
        """The one host read-back of a step."""
        return self.used > 0 and bool(self.words().cpu().any())


def _info_slot(words=1):
    """Where the next factorisation reports: a pool slot inside a pooled deferred region, else a fresh word."""
    return _POOL.next(words) if _POOL is not None else None


def _raise_if_not_pd(info):
    """LAPACK-style failure word -> NotPSDError.  Reading it synchronises the host with the stream; inside a
    `deferred_pd_checks()` region the words are collected instead and examined once by `.check()`."""
    if _DEFERRED is not None:
        _DEFERRED.append(info)
        return
    for i in info.reshape(-1).tolist():
        if i != 0:
            raise NotPSDError(f"Cholesky failed: leading minor {i} is not positive definite")


class deferred_pd_checks:
    """with deferred_pd_checks() as pending: ...; pending.check()

    Every Cholesky in the region runs optimistically (no read-back, so the host keeps issuing work ahead of the
    GPU); `check()` reads all failure words in one transfer and raises NotPSDError for the first failed
    factorisation.  Results computed after a failed factorisation are garbage (NaN) and must be discarded by the
    caller - training.ElboTrainer vetoes the update on the device and re-runs the step with immediate checks and
    jitter retries in that case.  With `pool` the words live in that InfoPool instead of separate tensors."""

    def __init__(self, pool=None):
        self.pool = pool

    def __enter__(self):
        global _DEFERRED, _POOL
        self._outer, self._outer_pool, self.infos = _DEFERRED, _POOL, []
        _DEFERRED, _POOL = self.infos, self.pool
        return self

    def __exit__(self, *exc):
        global _DEFERRED, _POOL
        _DEFERRED, _POOL = self._outer, self._outer_pool
        return False

    def check(self):
        if not self.infos:
            return
        words = torch.cat([i.reshape(1) for i in self.infos]).tolist()
        self.infos.clear()
        for k, w in enumerate(words):
            if w != 0:
                raise NotPSDError(f"Cholesky #{k} of the region failed: leading minor {w} is not positive definite")


class _CholeskyFn(Function):
    @staticmethod
    def forward(ctx, A, check):
        L = A.detach().clone(memory_format=torch.contiguous_format)
        info = ops.potrf_(L, info=_info_slot())
        if check:
            _raise_if_not_pd(info)
        L = L.tril_()
        ctx.save_for_backward(L)
        return L

    @staticmethod
    def backward(ctx, gL):
        (L,) = ctx.saved_tensors
        # Phi = tril(L^T gL) with halved diagonal ; gA = sym(L^{-T} Phi L^{-1})
        P = ops.gemm(L, gL.tril().contiguous(), transa=True)
        P = P.tril_()
        P.diagonal().mul_(0.5)
        X = ops.trsm_(L, P, trans=True)                 # L^{-T} Phi
        St = ops.trsm_(L, X.t().contiguous(), trans=True)  # (X L^{-1})^T
        return 0.5 * (St + St.t()), None


def cholesky(A, check=True):
    """Lower Cholesky factor (row-major); raises NotPSDError when `check`."""
    return _CholeskyFn.apply(A, check)


def psd_safe_cholesky(A, max_tries=3):
    """gpytorch.utils.cholesky.psd_safe_cholesky semantics: retry with
    jitter 1e-6*10^i (fp32) / 1e-8*10^i (fp64) on failure, then raise NotPSDError."""
    try:
        return cholesky(A)
    except NotPSDError:
        base = 1e-8 if A.dtype == torch.float64 else 1e-6
        eye = torch.eye(A.size(0), dtype=A.dtype, device=A.device)
        for i in range(max_tries):
            try:
                return cholesky(A + (base * 10 ** i) * eye)
            except NotPSDError:
                continue
        raise


class _TrsmFn(Function):
    @staticmethod
    def forward(ctx, L, B, trans):
        X = ops.trsm_(L, B.detach().clone(memory_format=torch.contiguous_format), trans=trans)
        ctx.trans = trans
        ctx.save_for_backward(L, X)
        return X

    @staticmethod
    def backward(ctx, gX):
        L, X = ctx.saved_tensors
        gB = ops.trsm_(L, gX.detach().clone(memory_format=torch.contiguous_format), trans=not ctx.trans)
        gL = None
        if ctx.needs_input_grad[0]:
            gL = ops.gemm(X, gB, transb=True, alpha=-1.0) if ctx.trans else ops.gemm(gB, X, transb=True, alpha=-1.0)
            gL = gL.tril_()
        return gL, gB, None


def trsm(L, B, trans=False):
    """Solve L X = B (trans=False) or L^T X = B (trans=True) for lower-triangular L."""
    vec = B.dim() == 1
    if vec:
        B = B.unsqueeze(-1)
    X = _TrsmFn.apply(L, B, trans)
    return X.squeeze(-1) if vec else X


def chol_solve(L, B):
    """(L L^T)^{-1} B."""
    return trsm(L, trsm(L, B, False), True)


class _LogdetFn(Function):
    @staticmethod
    def forward(ctx, L):
        ctx.save_for_backward(L)
        return ops.logdet_chol(L)

    @staticmethod
    def backward(ctx, g):
        (L,) = ctx.saved_tensors
        return torch.diag_embed(2.0 * g / L.diagonal())


class _GaussianTermsFn(Function):
    """(log|C|, e^T C^{-1} e, tr(C^{-1} G)) for symmetric positive definite C and symmetric G, with the
    explicit inverse C^{-1} = W^T W, W = chol(C)^{-1}: one factorisation, one triangular inversion and
    three products forward; two products backward (dC = g1 C^{-1} - g2 u u^T - g3 C^{-1} G C^{-1},
    u = C^{-1} e; de = 2 g2 u; dG = g3 C^{-1}).  Same quantities as the reference's Cholesky + solves
    (cmp_likelihood.py:60-78), an eighth of the dependent launches."""

    @staticmethod
    def forward(ctx, C, e, G):
        L = C.detach().clone(memory_format=torch.contiguous_format)
        info = ops.potrf_(L, info=_info_slot())
        _raise_if_not_pd(info)
        L = L.tril_()
        W = ops.trtri(L)
        Cinv = ops.gemm(W, W, transa=True)
        u = ops.gemm(Cinv, e.detach().reshape(-1, 1).contiguous()).reshape(-1)
        logdet = ops.logdet_chol(L)
        mean_term = (u * e.detach()).sum()
        covar_term = (Cinv * G.detach()).sum()
        ctx.save_for_backward(Cinv, u, G.detach())
        return logdet, mean_term, covar_term

    @staticmethod
    def backward(ctx, g1, g2, g3):
        Cinv, u, G = ctx.saved_tensors
        gC = ge = gG = None
        if ctx.needs_input_grad[0]:
            CG = ops.gemm(Cinv, G if G.is_contiguous() else G.contiguous())
            gC = ops.gemm(CG, Cinv, alpha=-1.0)
            gC = gC * g3 + g1 * Cinv - g2 * torch.outer(u, u)
        if ctx.needs_input_grad[1]:
            ge = 2.0 * g2 * u
        if ctx.needs_input_grad[2]:
            gG = g3 * Cinv
        return gC, ge, gG


def gaussian_terms(C, e, G, max_tries=3):
    """log|C|, e^T C^{-1} e and tr(C^{-1} G) with psd_safe_cholesky's jitter retries on C."""
    try:
        return _GaussianTermsFn.apply(C, e, G)
    except NotPSDError:
        base = 1e-8 if C.dtype == torch.float64 else 1e-6
        eye = torch.eye(C.size(0), dtype=C.dtype, device=C.device)
        for i in range(max_tries):
            try:
                return _GaussianTermsFn.apply(C + (base * 10 ** i) * eye, e, G)
            except NotPSDError:
                continue
        raise


class _CholInverseFn(Function):
    """W = chol(A)^{-1} (lower).  Backward entirely as products with W:
    dL = -tril(W^T gW W^T);  Phi = tril(L^T dL) with halved diagonal;  dA = sym(W^T Phi W)."""

    @staticmethod
    def forward(ctx, A):
        L = A.detach().clone(memory_format=torch.contiguous_format)
        info = ops.potrf_(L, info=_info_slot())
        _raise_if_not_pd(info)
        L = L.tril_()
        W = ops.trtri(L)
        ctx.save_for_backward(L, W)
        return W

    @staticmethod
    def backward(ctx, gW):
        L, W = ctx.saved_tensors
        T = ops.gemm(W, gW.tril().contiguous(), transa=True)          # W^T gW
        dL = ops.gemm(T, W, transb=True, alpha=-1.0).tril_()           # -(W^T gW) W^T
        P = ops.gemm(L, dL, transa=True).tril_()
        P.diagonal().mul_(0.5)
        S = ops.gemm(ops.gemm(W, P, transa=True), W)                   # W^T Phi W
        return 0.5 * (S + S.t())


def chol_inverse(A, max_tries=3):
    """Inverse Cholesky factor with psd_safe_cholesky's jitter retries."""
    try:
        return _CholInverseFn.apply(A)
    except NotPSDError:
        base = 1e-8 if A.dtype == torch.float64 else 1e-6
        eye = torch.eye(A.size(0), dtype=A.dtype, device=A.device)
        for i in range(max_tries):
            try:
                return _CholInverseFn.apply(A + (base * 10 ** i) * eye)
            except NotPSDError:
                continue
        raise


class _SplitMatmulFn(Function):
    """W @ K for a float64 W applied to float32 K at ~float32 accuracy on the TF32 tensor cores:
    W is carried as an FP32 pair (hi + lo, 48 significant bits), each half multiplied with 3xTF32.
    Backward: gK = W^T g (same pair), gW = tril(g K^T) (W is lower triangular) in float64."""

    @staticmethod
    def forward(ctx, W, K):
        Whi = W.detach().float()
        Wlo = (W.detach() - Whi.double()).float()
        K = K.detach().contiguous()
        Klo = ops.split_lo(K)
        out = ops.gemm_x3(Whi, K, Blo=Klo)
        ops.gemm_x3(Wlo, K, Blo=Klo, beta=1.0, C=out)
        ctx.save_for_backward(Whi, Wlo, K)
        return out

    @staticmethod
    def backward(ctx, g):
        Whi, Wlo, K = ctx.saved_tensors
        g = g.contiguous()
        gW = gK = None
        glo = ops.split_lo(g)
        if ctx.needs_input_grad[1]:
            gK = ops.gemm_x3(Whi, g, transa=True, Blo=glo)
            ops.gemm_x3(Wlo, g, transa=True, Blo=glo, beta=1.0, C=gK)
        if ctx.needs_input_grad[0]:
            gW = ops.gemm_x3(g, K, transb=True, Alo=glo).tril_().double()
        return gW, gK


def split_matmul(W, K):
    return _SplitMatmulFn.apply(W, K)


def logdet_from_chol(L):
    """log det(L L^T) = 2 sum log L_ii."""
    return _LogdetFn.apply(L)


# ---------------------------------------------------------------------- fused kernel-matrix ops
class _KernelSolveFn(Function):
    """W = (k(x,x) + c I)^{-1} B with a cached Cholesky factor and a low-rank adjoint."""

    @staticmethod
    def forward(ctx, meta, x, L, B, *hyp):
        # the factor's solve plan (built once by kernel_factor) serves this solve and its adjoint
        ctx.plan = getattr(L, "_dcgp_plan", None)
        W = ops.potrs_(L, B.detach().clone(memory_format=torch.contiguous_format), plan=ctx.plan)
        ctx.meta = meta
        ctx.save_for_backward(x, L, W, *hyp)
        return W

    @staticmethod
    def backward(ctx, gW):
        x, L, W, *hyp = ctx.saved_tensors
        meta = ctx.meta
        need = ctx.needs_input_grad
        gB = ops.potrs_(L, gW.detach().clone(memory_format=torch.contiguous_format), plan=ctx.plan)
        hg = [None] * len(hyp)
        if any(need[4:]):
            dLam = ops.gemm(gB, W, transb=True, alpha=-1.0)      # -(Lambda^{-1} gW) W^T
            dls, dos, _ = ops.gram_bwd(_spec(meta, hyp), x, None, dLam)
            hg = _hyper_grads(meta, hyp, dls, dos, need[4:])
        return (None, None, None, gB if need[3] else None, *hg)


def kernel_solve(meta, hyp, x, L, B):
    vec = B.dim() == 1
    if vec:
        B = B.unsqueeze(-1)
    W = _KernelSolveFn.apply(meta, x, L, B, *hyp)
    return W.squeeze(-1) if vec else W


def kernel_factor(meta, hyp, x, check=True):
    """Cholesky factor of k(x,x) + diag_const*I (no autograd: consumed by kernel_solve)."""
    with torch.no_grad():
        L = ops.gram(_spec(meta, hyp), x, None, diag_const=meta.diag_const)
        info = ops.potrf_(L, info=_info_slot())
        if check:
            _raise_if_not_pd(info)
        L._dcgp_plan = ops.solve_plan(L)
    return L


class _KernelMatmulFn(Function):
    """(k(x1,x2) [+ diag]) @ W; K is re-assembled in the backward instead of being saved."""

    @staticmethod
    def forward(ctx, meta, x1, x2, W, diag_dev, *hyp):
        out = ops.gram_matmul(_spec(meta, hyp), x1, x2, W.detach(), diag_dev=diag_dev, diag_const=meta.diag_const)
        ctx.meta, ctx.same = meta, x2 is None
        ctx.save_for_backward(x1, x2 if x2 is not None else x1, W, diag_dev if diag_dev is not None else W.new_zeros(1), *hyp)
        ctx.has_diag = diag_dev is not None
        return out

    @staticmethod
    def backward(ctx, gO):
        x1, x2, W, diag_dev, *hyp = ctx.saved_tensors
        meta, same = ctx.meta, ctx.same
        need = ctx.needs_input_grad
        gO = gO.contiguous()
        spec = _spec(meta, hyp)
        gW = None
        if need[3]:
            # K^T gO: the Gram matrix is symmetric when x2 is x1, otherwise k(x2, x1) is its transpose
            gW = ops.gram_matmul(spec, x1, None, gO, diag_dev=diag_dev if ctx.has_diag else None, diag_const=meta.diag_const) \
                if same else ops.gram_matmul(spec, x2, x1, gO)
        dx1 = dx2 = None
        hg = [None] * len(hyp)
        if need[1] or need[2] or any(need[5:]):
            dK = ops.gemm(gO, W, transb=True)
            dls, dos, dx1 = ops.gram_bwd(spec, x1, None if same else x2, dK, need_dx1=need[1])
            if not same and need[2]:
                dx2 = ops.gram_bwd(spec, x2, x1, dK.t().contiguous(), need_dx1=True)[2]
            hg = _hyper_grads(meta, hyp, dls, dos, need[5:])
        ddiag = (gO * W).sum().reshape(1) if (ctx.has_diag and need[4]) else None
        return (None, dx1, dx2, gW, ddiag, *hg)


def kernel_matmul(meta, hyp, x1, x2, W, diag_dev=None):
    vec = W.dim() == 1
    if vec:
        W = W.unsqueeze(-1)
    out = _KernelMatmulFn.apply(meta, x1, x2, W, diag_dev, *hyp)
    return out.squeeze(-1) if vec else out


# ---------------------------------------------------------------------- bag aggregation
class _GramBagsumFn(Function):
    @staticmethod
    def forward(ctx, meta, z, x, offsets, *hyp):
        out = ops.gram_bagsum(_spec(meta, hyp), z, x, offsets)
        ctx.meta = meta
        ctx.save_for_backward(z, x, offsets, *hyp)
        return out

    @staticmethod
    def backward(ctx, G):
        z, x, offsets, *hyp = ctx.saved_tensors
        need = ctx.needs_input_grad
        dls, dos, dz = ops.gram_bagsum_bwd(_spec(ctx.meta, hyp), z, x, offsets, G.contiguous(), need_dz=need[1])
        return (None, dz, None, None, *_hyper_grads(ctx.meta, hyp, dls, dos, need[4:]))


def gram_bagsum(meta, hyp, z, x, offsets):
    """out[a, i] = sum_{j in bag a} k(z_i, x_j)  (n_bags x M)."""
    return _GramBagsumFn.apply(meta, z, x, offsets, *hyp)


class _GramBagBlocksumFn(Function):
    @staticmethod
    def forward(ctx, meta, x, offsets, *hyp):
        out = ops.gram_bag_blocksum(_spec(meta, hyp), x, offsets)
        ctx.meta = meta
        ctx.save_for_backward(x, offsets, *hyp)
        return out

    @staticmethod
    def backward(ctx, g):
        x, offsets, *hyp = ctx.saved_tensors
        dls, dos = ops.gram_bag_blocksum_bwd(_spec(ctx.meta, hyp), x, offsets, g.contiguous())
        return (None, None, None, *_hyper_grads(ctx.meta, hyp, dls, dos, ctx.needs_input_grad[3:]))


def gram_bag_blocksum(meta, hyp, x, offsets):
    """out[a] = sum_{i,j in bag a} k(x_i, x_j)."""
    return _GramBagBlocksumFn.apply(meta, x, offsets, *hyp)


class _BagSumFn(Function):
    @staticmethod
    def forward(ctx, X, offsets):
        ctx.save_for_backward(offsets)
        ctx.n = X.size(0)
        return ops.bag_sum(X, offsets)

    @staticmethod
    def backward(ctx, g):
        (offsets,) = ctx.saved_tensors
        sizes = offsets[1:] - offsets[:-1]
        return torch.repeat_interleave(g, sizes, dim=0, output_size=ctx.n), None


def bag_sum(X, offsets):
    vec = X.dim() == 1
    if vec:
        X = X.unsqueeze(-1)
    out = _BagSumFn.apply(X, offsets)
    return out.squeeze(-1) if vec else out


class _WhitenedKLFn(Function):
    @staticmethod
    def forward(ctx, m, Ls):
        val, dm, dLs = ops.whitened_kl(m.detach(), Ls.detach(), need_grad=True)
        ctx.save_for_backward(dm, dLs)
        return val

    @staticmethod
    def backward(ctx, g):
        dm, dLs = ctx.saved_tensors
        return g * dm, g * dLs


class _VbaggEllFn(Function):
    @staticmethod
    def forward(ctx, z, S, T, offsets, noise):
        ell, dS, dT, dn = ops.vbagg_ell(z.detach(), S.detach(), T.detach(), offsets, noise.detach())
        ctx.save_for_backward(dS, dT, dn)
        ctx.noise_shape = noise.shape
        return ell

    @staticmethod
    def backward(ctx, g):
        dS, dT, dn = ctx.saved_tensors
        return None, g * dS, g * dT, None, (g * dn).reshape(ctx.noise_shape)


def vbagg_ell(z, S, T, offsets, noise):
    """sum_a E_q log N(z_a | mean_a f, noise / n_a) from the per-bag sums of the posterior mean (S) and
    covariance (T); bags are the contiguous row ranges `offsets`."""
    return _VbaggEllFn.apply(z, S, T, offsets, noise)


def whitened_kl(variational_mean, chol_variational_covar):
    """KL(N(m, tril(Ls) tril(Ls)^T) || N(0, I)) in one pass; the lower mask of
    CholeskyVariationalDistribution is applied inside the kernel."""
    return _WhitenedKLFn.apply(variational_mean, chol_variational_covar)


# ---------------------------------------------------------------------- whole VBagg ELBO (C composite)
class _VbaggElboFn(Function):
    """ELBO of one VBagg minibatch through dcgp_vbagg_elbo_fwd / _bwd (csrc/vbagg.cu): the forward call also runs the
    backward one (the gradients of a scalar do not depend on the upstream value except as a factor), so autograd's
    backward is a handful of scalar multiplications.  Inputs: Z, x, offsets, z, var_mean, chol_var, noise, mean_const,
    then the kernel hyper-parameter VALUES [ls_0, os_0, ...]."""

    @staticmethod
    def forward(ctx, meta, num_data, beta, Z, x, offsets, z, var_mean, chol_var, noise, mean_const, *hyp):
        spec = _spec(meta, hyp)
        need = any(ctx.needs_input_grad)
        elbo, g, info = ops.vbagg_elbo(spec, Z.detach(), x, offsets, z, var_mean.detach(), chol_var.detach(),
                                       noise.detach(), num_data, beta=beta,
                                       mean_const=None if mean_const is None else mean_const.detach().reshape(1),
                                       scale=1.0, need_grads=need, info=_info_slot())
        _raise_if_not_pd(info)
        ctx.meta, ctx.g = meta, g
        ctx.shapes = (noise.shape, None if mean_const is None else mean_const.shape, [h.shape for h in hyp])
        ctx.hyp_dtypes = [h.dtype for h in hyp]
        return elbo

    @staticmethod
    def backward(ctx, go):
        g, need = ctx.g, ctx.needs_input_grad
        noise_shape, mc_shape, hyp_shapes = ctx.shapes
        out = [None, None, None,
               go * g["Z"] if need[3] else None, None, None, None,
               go * g["var_mean"] if need[7] else None,
               go * g["chol_var"] if need[8] else None,
               (go * g["noise"]).reshape(noise_shape) if need[9] else None,
               (go * g["mean_const"]).reshape(mc_shape) if (mc_shape is not None and need[10]) else None]
        for c in range(len(ctx.meta.kinds)):
            ls_shape, os_shape = hyp_shapes[2 * c], hyp_shapes[2 * c + 1]
            n = 1
            for d in ls_shape:
                n *= d
            out.append((go * g["dls"][c, :n]).reshape(ls_shape) if need[11 + 2 * c] else None)
            out.append((go * g["dos"][c]).reshape(os_shape) if need[12 + 2 * c] else None)
        return tuple(out)


def vbagg_elbo(meta: KernelMeta, hyp, Z, x, offsets, z, var_mean, chol_var, noise, num_data, beta=1.0, mean_const=None):
    """Differentiable ELBO (not its negative) of a VBagg minibatch; see _VbaggElboFn."""
    return _VbaggElboFn.apply(meta, float(num_data), float(beta), Z, x, offsets, z, var_mean, chol_var, noise, mean_const,
                              *hyp)


# ---------------------------------------------------------------------- whole VariationalCMP ELBO (C composite)
class _CmpElboFn(Function):
    """ELBO of one VariationalCMP minibatch (with individuals noise) through dcgp_cmp_elbo_fwd / _bwd (csrc/cmp.cu).
    As in _VbaggElboFn the forward call runs both C calls and autograd's backward only scales the cached gradients.
    Inputs: Z, x, ext, y, z, var_mean, chol_var, noise, ind_noise, mean_const, then the hyper-parameter VALUES of the
    individuals kernel and of the bag kernel [ls_0, os_0, ...]."""

    @staticmethod
    def forward(ctx, meta_k, meta_l, lbda, num_data, beta, Z, x, ext, y, z, var_mean, chol_var, noise, ind_noise, mean_const,
                *hyp):
        nk = 2 * len(meta_k.kinds)
        hk, hl = hyp[:nk], hyp[nk:]
        need = any(ctx.needs_input_grad)
        elbo, g, info = ops.cmp_elbo(_spec(meta_k, hk), _spec(meta_l, hl), Z.detach(), x, ext, y, z, var_mean.detach(),
                                     chol_var.detach(), noise.detach(), ind_noise.detach(), lbda, num_data, beta=beta,
                                     mean_const=None if mean_const is None else mean_const.detach().reshape(1), scale=1.0,
                                     need_grads=need, info=_info_slot(3))
        _raise_if_not_pd(info)
        ctx.metas, ctx.g = (meta_k, meta_l), g
        ctx.shapes = (noise.shape, ind_noise.shape, None if mean_const is None else mean_const.shape, [h.shape for h in hyp])
        return elbo

    @staticmethod
    def backward(ctx, go):
        g, need = ctx.g, ctx.needs_input_grad
        noise_shape, ind_shape, mc_shape, hyp_shapes = ctx.shapes
        out = [None] * 5 + [go * g["Z"] if need[5] else None, None, None, None, None,
                            go * g["var_mean"] if need[10] else None, go * g["chol_var"] if need[11] else None,
                            (go * g["noise"]).reshape(noise_shape) if need[12] else None,
                            (go * g["ind_noise"]).reshape(ind_shape) if need[13] else None,
                            (go * g["mean_const"]).reshape(mc_shape) if (mc_shape is not None and need[14]) else None]
        pos = 15
        for meta, dls, dos in ((ctx.metas[0], g["dk_ls"], g["dk_os"]), (ctx.metas[1], g["dl_ls"], g["dl_os"])):
            for c in range(len(meta.kinds)):
                ls_shape, os_shape = hyp_shapes[pos - 15], hyp_shapes[pos - 14]
                n = 1
                for d in ls_shape:
                    n *= d
                out.append((go * dls[c, :n]).reshape(ls_shape) if need[pos] else None)
                out.append((go * dos[c]).reshape(os_shape) if need[pos + 1] else None)
                pos += 2
        return tuple(out)


def cmp_elbo(meta_k: KernelMeta, hyp_k, meta_l: KernelMeta, hyp_l, Z, x, ext, y, z, var_mean, chol_var, noise, ind_noise, lbda,
             num_data, beta=1.0, mean_const=None):
    """Differentiable ELBO (not its negative) of a VariationalCMP minibatch; see _CmpElboFn."""
    return _CmpElboFn.apply(meta_k, meta_l, float(lbda), float(num_data), float(beta), Z, x, ext, y, z, var_mean, chol_var,
                            noise, ind_noise, mean_const, *hyp_k, *hyp_l)
