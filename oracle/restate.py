"""TEST INFRASTRUCTURE ONLY — the oracle: a plain-PyTorch CPU restatement of the
reference's deconditional-GP hot path (Gram -> Cholesky/solves -> bag ELBO).

It is the checker for the CUDA path (tests/, __graft_entry__.smoke()) and the
`cpu_baseline` / `--impl reference` leg of bench.py.  Nothing under
deconditional_downscaling_b200/ imports it.

Every function cites the reference lines it follows.  gpytorch==1.4.0 (pinned
at /root/reference/requirements.txt:6, absent from this image) semantics are
restated from SURVEY.md Appendix A, always on the *Cholesky branch*.

Pinning: this file is checked in tests/test_oracle.py against
tests/golden/*.npz, which were produced by the reference's own `src/` code
executed over oracle/gpytorch_min (see tests/golden/make_golden.py), and
against independent closed forms (torch.distributions, torch.cdist).  The real
gpytorch could not be run, so parity with gpytorch's internals is UNPINNED.

dtype: whatever the inputs are (float64 for the parity gate, float32 to mimic
the reference's own precision incl. its float64 island for K_ZZ).
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
from torch.nn.functional import softplus

LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------
# kernel specification (mirrors ScaleKernel(RBF/Matern/RFF) [+ ...] trees)
# --------------------------------------------------------------------------
@dataclass
class Component:
    kind: str                       # 'rbf' | 'matern32' | 'rff'
    raw_lengthscale: torch.Tensor   # (nd,)
    raw_outputscale: torch.Tensor   # ()
    dims: Optional[Sequence[int]] = None   # active_dims (None = all columns)
    weights: Optional[torch.Tensor] = None  # (nd, D) RFF frequencies (buffer, not trained)

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale)

    @property
    def outputscale(self):
        return softplus(self.raw_outputscale)


@dataclass
class KernelSpec:
    components: List[Component] = field(default_factory=list)


def _select(x, dims):
    if x.dim() == 1:
        x = x.unsqueeze(-1)
    return x if dims is None else x[:, list(dims)]


def sq_dist(x1, x2, same):
    """gpytorch Kernel.covar_dist(square_dist=True): mean-centred quadratic
    expansion, zero diagonal if same & no grad, clamp_min(0).  (Appendix A
    `Squared distance`; used by every Gram on the path, e.g. src/models/cmp.py:60,65.)"""
    adj = x1.mean(-2, keepdim=True)
    x1, x2 = x1 - adj, x2 - adj
    n1 = x1.pow(2).sum(-1, keepdim=True)
    n2 = x2.pow(2).sum(-1, keepdim=True)
    res = torch.cat([-2.0 * x1, n1, torch.ones_like(n1)], -1) @ torch.cat([x2, torch.ones_like(n2), n2], -1).t()
    if same and not (x1.requires_grad or x2.requires_grad):
        res = res - torch.diag_embed(res.diagonal())
    return res.clamp_min(0)


def rff_features(c: Component, x):
    """src/kernels/rff_kernel.py:82-89 (`_featurize`): [cos(x W / l^T), sin(.)]."""
    xw = x @ (c.weights / c.lengthscale.reshape(-1, 1))
    return torch.cat([torch.cos(xw), torch.sin(xw)], -1)


def gram(spec: KernelSpec, x1, x2=None):
    """Dense k(x1, x2) for an additive tree of scaled base kernels.
    RBF / Matern per Appendix A; RFF per src/kernels/rff_kernel.py:91-114
    (x1==x2: Phi Phi^T / D, else (Phi1/D) Phi2^T)."""
    same = x2 is None or x2 is x1 or (x1.shape == x2.shape and torch.equal(x1, x2))
    if x2 is None:
        x2 = x1
    out = None
    for c in spec.components:
        a, b = _select(x1, c.dims), _select(x2, c.dims)
        if c.kind == "rbf":
            ls = c.lengthscale
            k = torch.exp(-0.5 * sq_dist(a / ls, b / ls, same))
        elif c.kind == "matern32":
            ls = c.lengthscale
            mean = a.mean(0, keepdim=True)
            r = sq_dist((a - mean) / ls, (b - mean) / ls, same).clamp_min(1e-30).sqrt()
            k = (1.0 + math.sqrt(3.0) * r) * torch.exp(-math.sqrt(3.0) * r)
        elif c.kind == "rff":
            D = c.weights.shape[1]
            k = rff_features(c, a) @ rff_features(c, b).t() / D
        else:
            raise ValueError(c.kind)
        k = k * c.outputscale
        out = k if out is None else out + k
    return out


def gram_diag(spec: KernelSpec, x):
    out = 0.0
    for c in spec.components:
        if c.kind == "rff":
            a = _select(x, c.dims)
            phi = rff_features(c, a)
            out = out + c.outputscale * (phi * phi).sum(-1) / c.weights.shape[1]
        else:
            out = out + c.outputscale * torch.ones(x.shape[0], dtype=x.dtype)
    return out


# --------------------------------------------------------------------------
# dense linear algebra helpers (Cholesky branch of gpytorch)
# --------------------------------------------------------------------------
def chol(A):
    """psd_safe_cholesky without the jitter retry (tests use PD inputs)."""
    return torch.linalg.cholesky(A)


def chol_solve(L, B):
    return torch.cholesky_solve(B, L)


def root_inv(A):
    """`A.root_inv_decomposition().root` on the Cholesky branch: R = L^{-T},
    R R^T = A^{-1} (Appendix A; src/models/cmp.py:67, variational_cmp.py:100)."""
    L = chol(A)
    eye = torch.eye(A.shape[0], dtype=A.dtype)
    return torch.linalg.solve_triangular(L, eye, upper=False).t()


def mvn_log_prob(mean, covar, value):
    """MultivariateNormal.log_prob (Appendix A)."""
    L = chol(covar)
    half = torch.linalg.solve_triangular(L, (value - mean).unsqueeze(-1), upper=False)
    return -0.5 * ((half * half).sum() + 2.0 * L.diagonal().log().sum() + mean.numel() * LOG2PI)


# --------------------------------------------------------------------------
# ExactCMP  
# --------------------------------------------------------------------------
def cmp_estimate_parameters(k, l, x, ext, lbda, noise_raw_outputscale=None, mu=None):
    """src/models/cmp.py:48-68: mu = mean(x), K = k(x,x) [+ sigma_ind^2 I],
    Lambda = l(ext,ext) + lbda*N*I, R = root of Lambda^{-1}."""
    N = x.shape[0]
    mu = torch.zeros(N, dtype=x.dtype) if mu is None else mu
    K = gram(k, x)
    if noise_raw_outputscale is not None:
        K = K + softplus(noise_raw_outputscale) * torch.eye(N, dtype=x.dtype)
    Lam = gram(l, ext) + lbda * N * torch.eye(N, dtype=x.dtype)
    return mu, K, root_inv(Lam)


def cme_aggregate_prior(l, ext, y, mu, K, R):
    """CMEAggregateMean (src/means/cme_aggregate_mean.py:35-55) and
    CMEAggregateKernel (src/kernels/cme_aggregate_kernel.py:37-70), literal:
    m = (B^T R)(R^T mu),  Q = (B^T R)(R^T K R)(R^T B),  B = l(ext, y)."""
    B = gram(l, ext, y)
    foo = B.t() @ R
    m = foo @ (R.t() @ mu)
    Q = foo @ (R.t() @ K @ R) @ foo.t()
    return m, Q


def exact_mll(m, Q, z, raw_noise):
    """ExactMarginalLogLikelihood over GaussianLikelihood.marginal
    (experiments/swiss_roll/models/exact_cmp.py:81,97; Appendix A):
    log N(z | m, Q + sigma^2 I) / n, sigma^2 = softplus(raw)+1e-4."""
    n = z.numel()
    noise = softplus(raw_noise).reshape(()) + 1e-4
    return mvn_log_prob(m, Q + noise * torch.eye(n, dtype=z.dtype), z) / n


def exact_cmp_posterior(k, l, x, ext, y, z, xs, mu, K, R, raw_noise, mu_s=None):
    """ExactCMP.predict (src/models/exact_cmp.py:126-155) +
    get_individuals_to_cmp_covar (src/models/cmp.py:87-104) +
    ExactCMPPredictionStrategy (cmp_prediction_strategy.py:47-63,101-114,141-178):
    C = (k(xs,x) R)(l(y,ext) R)^T ; mean = C alpha + mu*, alpha = (Q+s2 I)^{-1}(z-m);
    cov = k(xs,xs) - C (Q+s2 I)^{-1} C^T."""
    n = z.numel()
    m, Q = cme_aggregate_prior(l, ext, y, mu, K, R)
    noise = softplus(raw_noise).reshape(()) + 1e-4
    Lq = chol(Q + noise * torch.eye(n, dtype=z.dtype))
    alpha = chol_solve(Lq, (z - m).unsqueeze(-1)).squeeze(-1)
    C = (gram(k, xs, x) @ R) @ (gram(l, y, ext) @ R).t()
    mu_s = torch.zeros(xs.shape[0], dtype=xs.dtype) if mu_s is None else mu_s
    mean = C @ alpha + mu_s
    cov = gram(k, xs) - C @ chol_solve(Lq, C.t())
    return mean, cov


# --------------------------------------------------------------------------
# whitened SVGP q(f)  
# --------------------------------------------------------------------------
def masked_chol(chol_variational_covar):
    """CholeskyVariationalDistribution.forward: raw matrix times lower 0/1 mask."""
    return chol_variational_covar.tril()


def svgp_q(k, Z, x, var_mean, chol_var, mu_x=None, fp64_island=True):
    """src/variational/variational_strategy.py:14-68 (and gpytorch's own
    VariationalStrategy.forward used by src/models/variational_gp.py):
    L_Z = chol_fp64(K_ZZ + 1e-3 I); A~ = L_Z^{-1} K_Zx (fp64 -> input dtype);
    mu_q = A~^T m + mu(x); Sigma_q = K_xx + 1e-4 I + A~^T (L_S L_S^T - I) A~."""
    M, N = Z.shape[0], x.shape[0]
    full = gram(k, torch.cat([Z, x], 0))
    Kzz = full[:M, :M] + 1e-3 * torch.eye(M, dtype=x.dtype)
    Kzx = full[:M, M:]
    Kxx = gram(k, x)  # second kernel evaluation, variational_strategy.py:19-20
    wide = torch.float64 if fp64_island else x.dtype
    Lz = chol(Kzz.to(wide))
    At = torch.linalg.solve_triangular(Lz, Kzx.to(wide), upper=False).to(x.dtype)
    Ls = masked_chol(chol_var)
    mu_x = torch.zeros(N, dtype=x.dtype) if mu_x is None else mu_x
    mu_q = At.t() @ var_mean + mu_x
    S_minus_I = Ls @ Ls.t() - torch.eye(M, dtype=x.dtype)
    Sigma_q = Kxx + 1e-4 * torch.eye(N, dtype=x.dtype) + At.t() @ (S_minus_I @ At)
    return mu_q, Sigma_q


def whitened_kl(var_mean, chol_var):
    """variational_strategy.kl_divergence() = KL(N(m, L_S L_S^T) || N(0, I))
    (src/mlls/bag_variational_elbo.py:49; Appendix A `kl_divergence()`)."""
    Ls = masked_chol(chol_var)
    M = var_mean.numel()
    return 0.5 * ((Ls * Ls).sum() + (var_mean * var_mean).sum() - M - (Ls.diagonal() ** 2).log().sum())


# --------------------------------------------------------------------------
# CMP expected log-likelihood
# --------------------------------------------------------------------------
def elbo_computation_parameters(l, y, ext, lbda):
    """VariationalCMP.get_elbo_computation_parameters (src/models/variational_cmp.py:86-113)."""
    N = ext.shape[0]
    Lam = gram(l, ext) + lbda * N * torch.eye(N, dtype=ext.dtype)
    return root_inv(Lam), gram(l, y, ext)


def cmp_ell_with_noise(z, mu_q, Sigma_q, R, B, individuals_noise, raw_noise):
    """CMPLikelihood._compute_expected_log_prob_with_individuals_noise
    (src/likelihoods/cmp_likelihood.py:41-78), literal."""
    n = z.numel()
    noise = softplus(raw_noise).reshape(()) + 1e-4
    Lq = chol(Sigma_q)
    agg = B @ R
    ATA = agg @ (R.t() @ R) @ agg.t()
    C = individuals_noise * ATA + noise * torch.eye(n, dtype=z.dtype)
    Lc = chol(C)
    err = z - agg @ (R.t() @ mu_q)
    mean_term = (torch.linalg.solve_triangular(Lc, err.unsqueeze(-1), upper=False) ** 2).sum()
    buf = agg @ (R.t() @ Lq)
    covar_term = (torch.linalg.solve_triangular(Lc, buf, upper=False) ** 2).sum()
    logdet = 2.0 * Lc.diagonal().log().sum()
    return -0.5 * (n * LOG2PI + logdet + mean_term + covar_term)


def cmp_ell_with_noise_minimal(z, mu_q, Kxx_eff, At, chol_var, Lam_chol, B, individuals_noise, raw_noise):
    """Same quantity without R or chol(Sigma_q) (SURVEY §8a row 13 'minimal
    equivalent'): A = Lambda^{-1} B^T, H = A^T A, a = A^T mu_q, U = A~ A,
    G = A^T Kxx_eff A + (L_S^T U)^T (L_S^T U) - U^T U,  covar = tr(C^{-1} G).
    `Kxx_eff` = K_xx + 1e-4 I."""
    n = z.numel()
    noise = softplus(raw_noise).reshape(()) + 1e-4
    A = chol_solve(Lam_chol, B.t())
    H = A.t() @ A
    U = At @ A
    LsU = masked_chol(chol_var).t() @ U
    G = A.t() @ (Kxx_eff @ A) + LsU.t() @ LsU - U.t() @ U
    C = individuals_noise * H + noise * torch.eye(n, dtype=z.dtype)
    Lc = chol(C)
    err = z - A.t() @ mu_q
    mean_term = (torch.linalg.solve_triangular(Lc, err.unsqueeze(-1), upper=False) ** 2).sum()
    covar_term = torch.trace(chol_solve(Lc, G))
    logdet = 2.0 * Lc.diagonal().log().sum()
    return -0.5 * (n * LOG2PI + logdet + mean_term + covar_term)


def cmp_ell_without_noise(z, mu_q, var_q, R, B, raw_noise):
    """CMPLikelihood._compute_expected_log_prob_without_individuals_noise
    (src/likelihoods/cmp_likelihood.py:80-116): uses only diag(Sigma_q) (:99)."""
    n = z.numel()
    noise = softplus(raw_noise).reshape(()) + 1e-4
    agg = B @ R
    err = z - agg @ (R.t() @ mu_q)
    A = R @ agg.t()
    covar_term = (A * (var_q.unsqueeze(-1) * A)).sum()
    return -0.5 * (n * torch.log(2 * math.pi * noise) + ((err * err).sum() + covar_term) / noise)


# --------------------------------------------------------------------------
# VBagg expected log-likelihood
# --------------------------------------------------------------------------
def vbagg_ell(z, mu_q, Sigma_q, bag_sizes, raw_noise):
    """VBaggGaussianLikelihood.expected_log_prob
    (src/likelihoods/vbagg_gaussian_likelihood.py:78-111), literal incl. the
    per-bag block extraction (:54-76)."""
    noise = softplus(raw_noise).reshape(()) + 1e-4
    offs = [0]
    for b in bag_sizes:
        offs.append(offs[-1] + int(b))
    msum = torch.stack([mu_q[i:j].sum() for i, j in zip(offs[:-1], offs[1:])])
    bar = torch.stack([(Sigma_q[i:j, i:j] + torch.outer(mu_q[i:j], mu_q[i:j])).sum() for i, j in zip(offs[:-1], offs[1:])])
    sizes = torch.tensor([float(b) for b in bag_sizes], dtype=z.dtype)
    first = z.pow(2) - 2 * z * msum / sizes + bar / sizes.pow(2)
    first = first * (sizes / noise)
    second = torch.log(2 * math.pi * noise / sizes)
    return -0.5 * (first + second).sum()


# --------------------------------------------------------------------------
# ELBO combiner
# --------------------------------------------------------------------------
def bag_elbo(ell, kl, n_target, num_data, beta):
    """BagVariationalELBO.forward (src/mlls/bag_variational_elbo.py:47-69):
    ELL / target.size(0) - KL / (num_data / beta)."""
    return ell / n_target - kl / (num_data / beta)
