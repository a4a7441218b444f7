"""TEST INFRASTRUCTURE / CPU BASELINE ONLY — whole-step ports of the reference's training
steps on the CPU with plain PyTorch (autograd + torch.optim.Adam), built from the oracle
restatement (oracle/restate.py).  bench.py times these as `cpu_baseline` and as the
`--impl reference` arm (the reference itself needs gpytorch 1.4.0, which is absent and
un-installable offline — see oracle/gpytorch_min/__init__.py); tests use them as the
checker for the multi-GPU data-parallel step.  Never imported by the product package.

Each step follows the reference trainer line by line:
  * VariationalCMP: experiments/downscaling/models/variational_cmp.py:163-186
    (q = model(x); get_elbo_computation_parameters; -elbo; backward; Adam.step)
  * ExactCMP:       experiments/swiss_roll/models/exact_cmp.py:89-104
The linear algebra is the reference's (explicit root R = L^{-T} of Lambda^{-1}, chol(Sigma_q),
dense N_b x N_b q(f) covariance), associated the way gpytorch's lazy tensors evaluate it.
"""
import math

import torch
from torch.nn.functional import softplus

from . import restate as R


def make_vcmp_params(M, d, r, dtype=torch.float32, seed=0, additive_bag_kernel=True):
    """Raw (pre-softplus) parameters of a VariationalCMP with k = Scale(RBF ARD d) and
    l = Scale(Matern1.5 dims[0,1]) + Scale(RBF dims[2:]) — gpytorch initial values."""
    g = torch.Generator().manual_seed(seed)
    inv_sp1 = math.log(math.e - 1.0)  # softplus^-1(1): lengthscale 1, as the reference's builders set it
    p = {
        "Z": torch.randn(M, d, generator=g, dtype=dtype),
        "var_mean": torch.zeros(M, dtype=dtype),
        "chol_var": torch.eye(M, dtype=dtype),
        "k_raw_ls": torch.full((d,), inv_sp1, dtype=dtype),
        "k_raw_os": torch.zeros((), dtype=dtype),
        "raw_noise": torch.zeros(1, dtype=dtype),
        "noise_kernel_raw_os": torch.zeros((), dtype=dtype),
    }
    if additive_bag_kernel:
        p.update(l0_raw_ls=torch.full((2,), inv_sp1, dtype=dtype), l0_raw_os=torch.zeros((), dtype=dtype),
                 l1_raw_ls=torch.full((r - 2,), inv_sp1, dtype=dtype), l1_raw_os=torch.zeros((), dtype=dtype))
    else:
        p.update(l0_raw_ls=torch.full((r,), inv_sp1, dtype=dtype), l0_raw_os=torch.zeros((), dtype=dtype))
    return p


def vcmp_specs(p, r):
    k = R.KernelSpec([R.Component("rbf", p["k_raw_ls"], p["k_raw_os"], None)])
    if "l1_raw_ls" in p:
        l = R.KernelSpec([R.Component("matern32", p["l0_raw_ls"], p["l0_raw_os"], [0, 1]),
                          R.Component("rbf", p["l1_raw_ls"], p["l1_raw_os"], list(range(2, r)))])
    else:
        l = R.KernelSpec([R.Component("rbf", p["l0_raw_ls"], p["l0_raw_os"], None)])
    return k, l


def vcmp_loss(p, x, ext, y, z, lbda, num_data, beta=1.0, use_individuals_noise=True, detach_root=False):
    """-ELBO of one VariationalCMP minibatch, the reference's algorithm.  `detach_root` (tests only): no gradient
    through the root of Lambda^{-1}, which isolates the part of the bag-kernel gradient that flows through l(y, y~)."""
    k, l = vcmp_specs(p, ext.shape[1])
    mu_q, Sigma_q = R.svgp_q(k, p["Z"], x, p["var_mean"], p["chol_var"])
    Rt, B = R.elbo_computation_parameters(l, y, ext, lbda)
    if detach_root:
        Rt = Rt.detach()
    n = z.numel()
    noise = softplus(p["raw_noise"]).reshape(()) + 1e-4
    agg = B @ Rt
    if use_individuals_noise:
        Lq = R.chol(Sigma_q)
        ATA = (agg @ Rt.t()) @ (Rt @ agg.t())
        C = softplus(p["noise_kernel_raw_os"]) * ATA + noise * torch.eye(n, dtype=z.dtype)
        Lc = R.chol(C)
        err = z - agg @ (Rt.t() @ mu_q)
        mean_term = (torch.linalg.solve_triangular(Lc, err.unsqueeze(-1), upper=False) ** 2).sum()
        buf = agg @ (Rt.t() @ Lq)
        covar_term = (torch.linalg.solve_triangular(Lc, buf, upper=False) ** 2).sum()
        ell = -0.5 * (n * R.LOG2PI + 2.0 * Lc.diagonal().log().sum() + mean_term + covar_term)
    else:
        ell = R.cmp_ell_without_noise(z, mu_q, Sigma_q.diagonal(), Rt, B, p["raw_noise"])
    kl = R.whitened_kl(p["var_mean"], p["chol_var"])
    return -R.bag_elbo(ell, kl, n, num_data, beta)


class CpuVcmpTrainer:
    """Holds leaf parameters + Adam, runs reference-style steps on the CPU."""

    def __init__(self, params, lr=0.01, use_individuals_noise=True):
        self.p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        self.use_individuals_noise = use_individuals_noise
        self.opt = torch.optim.Adam(list(self.p.values()), lr=lr)

    def step(self, x, ext, y, z, lbda, num_data, beta=1.0):
        self.opt.zero_grad()
        loss = vcmp_loss(self.p, x, ext, y, z, lbda, num_data, beta, self.use_individuals_noise)
        loss.backward()
        self.opt.step()
        return loss.detach()


def exact_cmp_loss(k, l, raw_noise, x, ext, y, z, lbda, detach_root=False):
    """-MLL of the ExactCMP fit step (experiments/swiss_roll/models/exact_cmp.py:94-97)."""
    mu, K, Rt = R.cmp_estimate_parameters(k, l, x, ext, lbda)
    if detach_root:
        Rt = Rt.detach()
    m, Q = R.cme_aggregate_prior(l, ext, y, mu, K, Rt)
    return -R.exact_mll(m, Q, z, raw_noise)


def exact_cmp_phases(x, ext, y, z, xs, lbda, dtype=torch.float64):
    """Gram + Cholesky + solves + posterior of the ExactCMP sweep (BASELINE config 4) on the CPU,
    minimal-flop form so the CPU baseline is not handicapped: K, Lambda, potrf, potrs, K W, Q,
    MLL, posterior mean / variance on xs."""
    d, r = x.shape[1], ext.shape[1]
    inv_sp1 = math.log(math.e - 1.0)
    k = R.KernelSpec([R.Component("rbf", torch.full((d,), inv_sp1, dtype=dtype), torch.zeros((), dtype=dtype), None)])
    l = R.KernelSpec([R.Component("rbf", torch.full((r,), inv_sp1, dtype=dtype), torch.zeros((), dtype=dtype), None)])
    N, n = x.shape[0], y.shape[0]
    with torch.no_grad():
        K = R.gram(k, x)
        Lam = R.gram(l, ext)
        Lam.diagonal().add_(lbda * N)
        L = torch.linalg.cholesky(Lam)
        W = torch.cholesky_solve(R.gram(l, ext, y), L)
        KW = K @ W
        Q = W.t() @ KW
        noise = softplus(torch.zeros((), dtype=dtype)) + 1e-4
        Lq = torch.linalg.cholesky(Q + noise * torch.eye(n, dtype=dtype))
        alpha = torch.cholesky_solve(z.unsqueeze(-1), Lq).squeeze(-1)
        mll = -0.5 * ((z * alpha).sum() + 2 * Lq.diagonal().log().sum() + n * R.LOG2PI) / n
        C = R.gram(k, xs, x) @ W
        mean = C @ alpha
        V = torch.linalg.solve_triangular(Lq, C.t(), upper=False)
        var = R.gram_diag(k, xs) - (V * V).sum(0)
    return {"mll": mll, "post_mean": mean, "post_var": var, "logdet": 2 * Lq.diagonal().log().sum()}
