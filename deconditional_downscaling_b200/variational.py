"""Whitened sparse variational GP machinery with gpytorch's class surface
(CholeskyVariationalDistribution, VariationalStrategy) and the reference's override
(src/variational/variational_strategy.py:10-68), evaluated lazily: q(f) is a
VariationalPosterior that hands the likelihoods exactly the reductions they need
(A^T Sigma_q A, diag Sigma_q, per-bag sums) instead of a dense N x N covariance.

FP64 island kept as in the reference (variational_strategy.py:35-44): K_ZZ + 1e-3 I is
factored and K_Zx solved in double even when the model runs in float32.
"""
import torch
from torch import nn

from . import functional as F
from .distributions import MultivariateNormal
from .kernels import _as_2d
from .lazy import LazyCovariance
from .module import Module


import contextlib
import os

# Overlap of the K_ZZ factor (a 1 ms chain of single-CTA kernels) with the caller's next work on a side stream.
# DCGP_SIDE_STREAM = "auto" (default): on inside `side_stream_scope()`, which the CUDA-graph trainer enters - a captured
# step has a fixed schedule and gains 7 % (B200, bench workload: 13.65 -> 12.7 ms, six runs, never slower);
# "1": always on (eager steps: 15.9 -> 15.2 ms when it works, but some eager runs were bimodal at 21 ms: two serial
# chains and a persistent GEMM grid compete for the SMs left free); "0": never.
SIDE_STREAM = os.environ.get("DCGP_SIDE_STREAM", "auto")
_SCOPE = [0]
# prefetch() also whitens K_Zx and forms D on the side stream (B200, captured bench step: 12.58 -> 12.20 ms, 3 of 3 runs)
SIDE_WHITEN = os.environ.get("DCGP_SIDE_WHITEN", "1") == "1"
# FP32 models whiten K_Zx with L_Z^{-1} as an FP32 hi / lo pair on 3xTF32 products (functional.split_matmul).  The rows of
# L_Z^{-1} are long alternating sums (|L_Z^{-1}| ~ cond(K_ZZ + 1e-3 I)^(1/2) ~ 30 - 100), so the ~5e-7 product error
# becomes ~2e-4 on A~: q(f).mean of the config-5 minibatch is 2.8e-4 off the FP64 oracle (inside the 1e-3 gate) and
# the gradients of Z and of the variational mean, which feel the residual z - A^T mu_q, 3.5e-3.  WHITEN64 ("DCGP_WHITEN64=1")
# evaluates K_Zx and the whitening product in float64 instead: 3e-6 / 6e-4 / 1.6e-3, at 1.6 ms per step (92.8 -> 80.7
# steps/s on B200) - an accuracy option, not the default (profiles/r02_parity.txt).  Adding only the image of the part
# of K_Zx that FP32 storage drops was tried and changes nothing: the product, not the storage, is where the error is.
WHITEN64 = os.environ.get("DCGP_WHITEN64", "0") == "1"


def side_stream_active():
    if isinstance(SIDE_STREAM, bool):
        return SIDE_STREAM
    if SIDE_STREAM in ("0", "1"):
        return SIDE_STREAM == "1"
    return _SCOPE[0] > 0


@contextlib.contextmanager
def side_stream_scope():
    _SCOPE[0] += 1
    try:
        yield
    finally:
        _SCOPE[0] -= 1


_SIDE = {}


def _quiet_accumulate_grad_warning():
    """The gradients of Z and of the kernel hyper-parameters arrive from two streams by design; autograd orders them
    correctly and only warns about the extra synchronisation."""
    fn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
    if fn is not None and not _SIDE.get("_quiet"):
        fn(False)
        _SIDE["_quiet"] = True


class CholeskyVariationalDistribution(Module):
    def __init__(self, num_inducing_points, batch_shape=None, mean_init_std=1e-3, **kwargs):
        super().__init__()
        self.num_inducing_points = num_inducing_points
        self.mean_init_std = mean_init_std
        self.register_parameter("variational_mean", nn.Parameter(torch.zeros(num_inducing_points)))
        self.register_parameter("chol_variational_covar", nn.Parameter(torch.eye(num_inducing_points)))

    def initialize_variational_distribution(self):
        """First-call init from the whitened prior N(0, I): mean <- 1e-3 * randn, chol <- I
        (`_VariationalStrategy.__call__`)."""
        with torch.no_grad():
            self.variational_mean.zero_()
            self.variational_mean.add_(torch.randn_like(self.variational_mean), alpha=self.mean_init_std)
            self.chol_variational_covar.copy_(torch.eye(self.num_inducing_points).to(self.chol_variational_covar))


class VariationalPosterior(MultivariateNormal):
    """q(f) at inputs x:  mu_q = A~^T m + mu(x),  Sigma_q = K_xx + 1e-4 I + A~^T (L_S L_S^T - I) A~,
    A~ = L_Z^{-1} K_Zx,  L_Z = chol(K_ZZ + 1e-3 I)   (src/variational/variational_strategy.py:14-68).
    Everything is computed on demand and cached for the lifetime of the object (one step)."""

    def __init__(self, kernel, mean_module, Z, x, variational_mean, chol_variational_covar):
        self.kernel, self.mean_module = kernel, mean_module
        self.Z, self.x = Z, _as_2d(x)
        self.m, self.Ls_raw = variational_mean, chol_variational_covar
        self.kops = kernel.ops(self.x)
        self._c = {}
        n = self.x.size(0)
        self._covar = LazyCovariance(self._dense_covar, self._variance, n)

    # ---- cached pieces -------------------------------------------------------------------
    def _get(self, key, fn):
        self._join_side()
        if key not in self._c:
            self._c[key] = fn()
        return self._c[key]

    def _join_side(self):
        """First use of anything after prefetch(): order the consumer stream after the side stream."""
        side = self._c.pop("_side", None)
        if side is not None:
            main = torch.cuda.current_stream(self.x.device)
            main.wait_stream(side)
            for key in self._c.pop("_side_keys", ()):
                v = self._c.get(key)
                if torch.is_tensor(v):
                    v.record_stream(main)

    @property
    def dtype(self):
        return self.x.dtype

    @property
    def Ls(self):
        return self._get("Ls", lambda: self.Ls_raw.to(self.dtype).tril())

    def _Kzz64(self):
        """K_ZZ + 1e-3 I of the float64 island (variational_strategy.py:35-44).  The reference evaluates the kernel in
        the model dtype and casts the matrix; here the M x M matrix is EVALUATED in double from the (cast) inducing
        points and hyper-parameters: its factor amplifies entry errors by cond(K_ZZ + 1e-3 I) ~ 1e5, which took the
        gradients of Z and of the variational mean of FP32 models 3e-3 off the FP64 oracle
        (tests/test_parity_scale_gpu.py); the cost is M^2 double-precision kernel evaluations."""
        Zd = self.Z.double()
        return self.kernel.ops(Zd).gram(Zd, None, diag_const=1e-3)

    @property
    def Lz(self):
        """chol(K_ZZ + 1e-3 I) in float64 (psd_safe_cholesky semantics)."""
        return self._get("Lz", lambda: F.psd_safe_cholesky(self._Kzz64()))

    @property
    def Wz(self):
        """L_Z^{-1} in float64 (FP32 models): the whitening products then run as 3xTF32 tensor-core GEMMs on
        an FP32 hi/lo pair of it instead of float64 triangular solves against 10^3 .. 10^4 columns."""
        return self._get("Wz", lambda: F.chol_inverse(self._Kzz64()))

    def prefetch(self):
        """Start the K_ZZ factorisation + inversion on a side stream.  It is a chain of single-CTA kernels (about
        1 ms at M = 1024) that is independent of the Lambda factorisation the caller issues next on the main
        stream (VariationalCMP) - both are latency-bound, together they cost the longer of the two.  The autograd
        engine runs the backward of these nodes on the same side stream."""
        if self.dtype == torch.float64 or not self.x.is_cuda or "Wz" in self._c or not side_stream_active():
            return
        dev = self.x.device
        side = _SIDE.get(dev)
        if side is None:
            side = _SIDE[dev] = torch.cuda.Stream(device=dev)
        _quiet_accumulate_grad_warning()
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            keys = ["Wz"]
            self._c["Wz"] = F.chol_inverse(self._Kzz64())
            if SIDE_WHITEN:
                # ... and what follows from it without the CME solve: A~ = L_Z^{-1} K_Zx and D = L_S L_S^T - I.  Their
                # backward nodes run on the side stream as well, next to the second big solve of the backward pass.
                self._c["At"] = self._whiten_kzx()
                self.D
                keys += ["At", "Ls", "D"]
        self._c["_side"], self._c["_side_keys"] = side, keys

    def _whiten(self, K):
        """L_Z^{-1} K for K (M x cols) in the model dtype."""
        if self.dtype == torch.float64:
            return F.trsm(self.Lz, K, trans=False)
        return F.split_matmul(self.Wz, K)

    def _whiten_kzx(self):
        """A~ = L_Z^{-1} K_Zx for the model's own x (see WHITEN64)."""
        if WHITEN64 and self.dtype != torch.float64:
            Zd, xd = self.Z.double(), self.x.double()
            return F.matmul(self.Wz, self.kernel.ops(Zd).gram(Zd, xd)).to(self.dtype)
        return self._whiten(self.kops.gram(self.Z, self.x))

    @property
    def At(self):
        """A~ = L_Z^{-1} K_Zx (float64 factor, result in the model dtype), M x N."""
        return self._get("At", self._whiten_kzx)

    @property
    def loc(self):
        return self._get("mean", lambda: F.matmul(self.At, self.m.to(self.dtype), transa=True) + self.mean_module(self.x).to(self.dtype))

    @loc.setter
    def loc(self, v):
        pass

    @property
    def D(self):
        """L_S L_S^T - I (M x M), the middle matrix of Sigma_q = K_xx + 1e-4 I + A~^T (L_S L_S^T - I) A~
        (variational_strategy.py:57-66).  Everything below is written on D - never as the difference of the
        two large quadratic forms |L_S^T A~|^2 - |A~|^2, which cancels (L_S starts at I) and would expose the
        truncation bias of single-pass TF32 products; D itself is formed with 3xTF32."""
        def fn():
            Ls = self.Ls
            eye = torch.eye(Ls.size(0), dtype=Ls.dtype, device=Ls.device)
            return F.matmul(Ls, Ls, transb=True, precise=True) - eye
        return self._get("D", fn)

    def _variance(self):
        def fn():
            At = self.At
            return self.kops.diag(self.x) + 1e-4 + (F.matmul(self.D, At) * At).sum(0)
        return self._get("var", fn)

    def _dense_covar(self):
        def fn():
            At = self.At
            return self.kops.gram(self.x, None, diag_const=1e-4) + F.matmul(At, F.matmul(self.D, At), transa=True)
        return self._get("cov", fn)

    # ---- reductions for the likelihoods ---------------------------------------------------------
    def quad_form(self, A):
        """A^T Sigma_q A for A (N x n) without forming Sigma_q:
        A^T (K_xx + 1e-4 I) A + U^T D U,  U = A~ A  (SURVEY.md §8a row 13)."""
        KA = self.kops.matmul(self.x, None, A, diag_const=1e-4)
        U = F.matmul(self.At, A)
        return F.matmul(A, KA, transa=True) + F.matmul(U, F.matmul(self.D, U), transa=True)

    def bag_stats(self, offsets):
        """S_a = 1_a^T mu_q and T_a = 1_a^T Sigma_q 1_a for contiguous bags, from the bag-summed
        cross Gram (K_Zx is never formed): u_a = L_Z^{-1} (K_Zx 1_a),
        T_a = 1_a^T K_xx 1_a + 1e-4 n_a + u_a^T D u_a  (SURVEY.md §8a row 15)."""
        sizes = (offsets[1:] - offsets[:-1]).to(self.dtype)
        KbZ = self.kops.bagsum(self.Z, self.x, offsets)                    # n_bags x M
        Ub = self._whiten(KbZ.t().contiguous())                               # M x n_bags
        S = F.matmul(Ub, self.m.to(self.dtype), transa=True) + F.bag_sum(self.mean_module(self.x).to(self.dtype), offsets)
        T = self.kops.bag_blocksum(self.x, offsets) + 1e-4 * sizes + (F.matmul(self.D, Ub) * Ub).sum(0)
        return S, T


class VariationalStrategy(Module):
    def __init__(self, model, inducing_points, variational_distribution, learn_inducing_locations=True):
        super().__init__()
        object.__setattr__(self, "model", model)
        inducing_points = inducing_points.clone()
        if inducing_points.dim() == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        if learn_inducing_locations:
            self.register_parameter("inducing_points", nn.Parameter(inducing_points))
        else:
            self.register_buffer("inducing_points", inducing_points)
        self._variational_distribution = variational_distribution
        self.register_buffer("variational_params_initialized", torch.tensor(0))
        self.register_buffer("updated_strategy", torch.tensor(True))

    def kl_divergence(self):
        """KL(N(m, L_S L_S^T) || N(0, I)) in one pass over L_S (bag_variational_elbo.py:49)."""
        vd = self._variational_distribution
        return F.whitened_kl(vd.variational_mean, vd.chol_variational_covar)

    def ensure_initialized(self):
        """First-call initialisation of q(u) (gpytorch _VariationalStrategy.__call__).  The flag is a device buffer
        (it is part of the checkpoint); once seen set it is remembered on the host so that steady-state steps do
        not synchronise on it."""
        if getattr(self, "_initialized_host", False):
            return
        if not bool(self.variational_params_initialized.item()):
            self._variational_distribution.initialize_variational_distribution()
            self.variational_params_initialized.fill_(1)
        self._initialized_host = True

    def _prior_modules(self):
        m = self.model
        if hasattr(m, "individuals_kernel"):       # VariationalCMP (src/models/variational_cmp.py:76-79)
            return m.individuals_kernel, m.individuals_mean
        return m.covar_module, m.mean_module        # VariationalGP (src/models/variational_gp.py:54-55)

    def __call__(self, x, prior=False, **kwargs):
        if prior:
            return self.model.forward(x, **kwargs)
        self.ensure_initialized()
        kernel, mean = self._prior_modules()
        vd = self._variational_distribution
        Z = self.inducing_points
        x = _as_2d(x)
        post = VariationalPosterior(kernel, mean, Z.to(x.dtype), x, vd.variational_mean, vd.chol_variational_covar)
        post.prefetch()
        return post
