"""ExactCMP fit + posterior as explicit phases on libdcgp kernels (no autograd): the
"Gram + Cholesky + solve" composite of BASELINE.json's config 4.

Same mathematics as ExactCMP.forward / predict (src/models/exact_cmp.py:107-155,
src/kernels/cme_aggregate_kernel.py:37-70, cmp_prediction_strategy.py:47-63,141-178) in the
minimal form: W = Lambda^{-1} l(y~, y), Q = W^T (K W), alpha = (Q + s2 I)^{-1} z,
mean* = k(x*, x) W alpha, var* = k** - |Lq^{-1} (k(x*, x) W)^T|^2.
"""
import math
from contextlib import contextmanager

import torch

from . import ops


class PhaseTimer:
    """CUDA-event timer on the current stream; collects seconds per named phase."""

    def __init__(self, enabled=True):
        self.enabled, self.events = enabled, []

    @contextmanager
    def phase(self, name):
        if not self.enabled:
            yield
            return
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        self.events.append((name, s, e))

    def seconds(self):
        torch.cuda.synchronize()
        out = {}
        for name, s, e in self.events:
            out[name] = out.get(name, 0.0) + s.elapsed_time(e) * 1e-3
        return out


def rbf_spec(d, dtype, device, lengthscale=1.0, outputscale=math.log(2.0)):
    """Scale(RBF ARD) at gpytorch's initial values: lengthscale 1, outputscale softplus(0)."""
    return ops.FlatSpec(["rbf"], [list(range(d))], [torch.full((d,), lengthscale, dtype=dtype, device=device)],
                        [torch.full((1,), outputscale, dtype=dtype, device=device)])


def padded(rows, cols, dtype, device, mult=4):
    """rows x cols view of a buffer whose row stride is rounded up to `mult` elements (16 B in FP32): the tall-skinny
    N x n operands here have n = number of bags, rarely a multiple of 4, and TMA (tcgen05 GEMM) needs 16-byte
    aligned rows - an unaligned W sends K W to the scalar-load mma.sync kernel (measured at N = 32768,
    n = 327, FP32: 13.8 ms = 51 TF/s)."""
    if dtype != torch.float32:      # FP64 (DMMA kernel, cp.async loads): measured no gain from aligned rows
        mult = 1
    ld = -(-cols // mult) * mult
    return torch.empty(rows, ld, dtype=dtype, device=device)[:, :cols]


def exact_cmp_phases(x, ext, y, z, xs, lbda, timer=None, k_spec=None, l_spec=None, noise=math.log(2.0) + 1e-4):
    """Runs Gram(K), Gram(Lambda), potrf, potrs, K W, Q / MLL and the posterior on xs.
    Returns dict(mll, logdet, post_mean, post_var, info)."""
    timer = timer or PhaseTimer(False)
    dt, dev = x.dtype, x.device
    N, n = x.size(0), y.size(0)
    if ext.dim() == 1:
        ext = ext.unsqueeze(-1)
    if y.dim() == 1:
        y = y.unsqueeze(-1)
    k_spec = k_spec or rbf_spec(x.size(1), dt, dev)
    l_spec = l_spec or rbf_spec(ext.size(1), dt, dev)
    with timer.phase("gram_K"):
        # the materialised Gram build of SURVEY 8d.4 (HBM-write roofline); nothing below reads it: K W and the posterior
        # cross-covariance go through the fused Gram-times-matrix entry point, which never holds K
        K = ops.gram(k_spec, x, None)
    del K
    with timer.phase("gram_Lambda"):
        Lam = ops.gram(l_spec, ext, None, diag_const=lbda * N)
    with timer.phase("potrf"):
        info = ops.potrf_(Lam)
    with timer.phase("gram_B"):
        W = ops.gram(l_spec, ext, y, out=padded(N, n, dt, dev))   # l(y~, y), N x n
    with timer.phase("potrs"):
        ops.potrs_(Lam, W)                              # W = Lambda^{-1} l(y~, y)
    with timer.phase("KW"):
        KW = ops.gram_matmul(k_spec, x, None, W)        # FP32: tcgen05 kernel generating the Gram tiles in shared memory
    with timer.phase("Q_mll"):
        Q = ops.gemm(W, KW, transa=True)
        Q.diagonal().add_(noise)
        info_q = ops.potrf_(Q)
        Lq = Q.tril_()
        alpha = z.reshape(-1, 1).clone()
        ops.potrs_(Lq, alpha)
        logdet = ops.logdet_chol(Lq)
        mll = -0.5 * ((z * alpha.squeeze(-1)).sum() + logdet + n * math.log(2 * math.pi)) / n
    with timer.phase("posterior"):
        C = ops.gram_matmul(k_spec, xs, x, W)            # k(x*, x) W, N* x n
        mean = ops.gemm(C, alpha).squeeze(-1)
        Vt = C.t().contiguous()
        ops.trsm_(Lq, Vt, trans=False)                  # Lq^{-1} C^T
        var = k_spec.outputscales[0].sum() - (Vt * Vt).sum(0)
    return {"mll": mll, "logdet": logdet, "post_mean": mean, "post_var": var, "info": info, "info_q": info_q}


def exact_cmp_flops(N, n, d, n_star):
    """Algorithmic flops of the composite: N^3/3 + 6 N^2 n + Gram cross terms."""
    return {
        "potrf": N ** 3 / 3.0,
        "potrs": 4.0 * N * N * n,
        "KW": 2.0 * N * N * n,
        "Q_mll": 2.0 * N * n * n,
        "gram": 2.0 * (d + 2) * N * N * 2,
        "posterior": 2.0 * n_star * N * n,
    }
