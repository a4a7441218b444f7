"""GPU tests at (or towards) BASELINE.json's full sizes through size-independent properties, plus
mid-scale parity of the fused VBagg path and of the row-block sharded Gram product against the oracle.
"""
import math

import pytest
import torch
from torch.nn.functional import softplus

from oracle import restate as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ops():
    from deconditional_downscaling_b200 import ops as o
    return o


def rbf(d, dtype):
    from deconditional_downscaling_b200.exact import rbf_spec
    return rbf_spec(d, dtype, DEV)


@pytest.mark.parametrize("dtype,n", [(torch.float64, 8192), (torch.float32, 8192), (torch.float64, 16384)])
def test_cholesky_reconstruction_and_solve_residual_full_size(dtype, n):
    """L L^T = A and A X = B at sweep sizes (config 4); FP64 to 1e-12, FP32 (3xTF32) to 1e-4."""
    o = ops()
    torch.manual_seed(n)
    x = torch.randn(n, 3, dtype=dtype, device=DEV)
    A = o.gram(rbf(3, dtype), x, None, diag_const=0.01 * n)
    assert torch.equal(A, A.t())                               # Gram symmetry is exact (difference form)
    L = A.clone()
    info = o.potrf_(L)
    assert info.item() == 0
    L = L.tril_()
    m = n // 100
    B = torch.randn(n, m, dtype=dtype, device=DEV)
    X = o.potrs_(L, B.clone())
    resid = (A.double() @ X.double() - B.double()).norm() / B.double().norm()
    assert resid.item() < (1e-12 if dtype == torch.float64 else 1e-4)
    # reconstruction on a random probe (avoids forming L L^T): A v = L (L^T v)
    v = torch.randn(n, 8, dtype=torch.float64, device=DEV)
    rec = L.double() @ (L.double().t() @ v)
    assert ((rec - A.double() @ v).norm() / (A.double() @ v).norm()).item() < (1e-13 if dtype == torch.float64 else 1e-5)
    # log-det from the factor equals the sum of log pivots and is finite
    ld = o.logdet_chol(L)
    assert math.isfinite(ld.item())
    assert abs(ld.item() - 2 * L.diagonal().double().log().sum().item()) < 1e-6 * abs(ld.item())


def test_gemm_linearity_full_size():
    """C(a A1 + A2, B) = a C(A1, B) + C(A2, B) on the tcgen05 path (FP32 8192-class shapes)."""
    o = ops()
    torch.manual_seed(1)
    A1, A2 = torch.randn(4096, 2048, device=DEV), torch.randn(4096, 2048, device=DEV)
    B = torch.randn(3072, 2048, device=DEV)
    o.TF32_MODE = "fast"
    try:
        lhs = o.gemm(2.0 * A1 + A2, B, transb=True)
        rhs = 2.0 * o.gemm(A1, B, transb=True) + o.gemm(A2, B, transb=True)
    finally:
        o.TF32_MODE = "auto"
    assert ((lhs - rhs).norm() / rhs.norm()).item() < 2e-3


@pytest.mark.parametrize("transa,transb", [(False, False), (True, False), (False, True)])
def test_large_products_with_unaligned_rows_take_aligned_copies(transa, transb):
    """An N x 327 FP32 operand (rows not 16-byte aligned, as W = Lambda^{-1} l(y~, y) with 327 bags) in a product large
    enough for the tcgen05 kernel: ops.gemm re-lays the operand on aligned rows; result vs FP64 at TF32 tolerance,
    operands untouched, result contiguous."""
    o = ops()
    torch.manual_seed(2)
    N, n = 8192, 327
    K = torch.randn(N, N, device=DEV)
    W = torch.randn(N, n, device=DEV)
    A = W if transa else K                       # op(A) is n x N (transa) or N x N
    B = torch.randn(n, N, device=DEV)[:, :N] if transb else W
    if transb:                                   # op(B) = B^T: N x n from an n x N matrix with aligned rows; unaligned A instead
        A = torch.randn(N, N + 1, device=DEV)[:, :N] if not transa else W
    W0 = W.clone()
    C = o.gemm(A, B, transa=transa, transb=transb)
    a = A.double().t() if transa else A.double()
    b = B.double().t() if transb else B.double()
    ref = a @ b
    assert C.is_contiguous() and C.shape == ref.shape
    assert ((C.double() - ref).norm() / ref.norm()).item() < 2e-3
    assert torch.equal(W, W0)
    assert o._tma_rows(W) is not W and o._tma_rows(K) is K
    assert (o._ld(o._tma_rows(W)) & 3) == 0 and torch.equal(o._tma_rows(W), W)


def test_exact_cmp_composite_matches_oracle_port():
    """Gram + Cholesky + solve + posterior composite (config 4 shape, N = 3000) vs the CPU port, FP64 1e-6."""
    from deconditional_downscaling_b200 import exact
    from oracle import port
    dt = torch.float64
    g = torch.Generator().manual_seed(3)
    N, n = 3000, 30
    x = torch.randn(N, 3, generator=g, dtype=dt)
    y = torch.randn(n, 3, generator=g, dtype=dt)
    ext = y[torch.arange(N) * n // N] + 0.3 * torch.randn(N, 3, generator=g, dtype=dt)
    z = torch.randn(n, generator=g, dtype=dt)
    ref = port.exact_cmp_phases(x, ext, y, z, x[:500], 0.01)
    got = exact.exact_cmp_phases(x.to(DEV), ext.to(DEV), y.to(DEV), z.to(DEV), x[:500].to(DEV), 0.01)
    for key in ("mll", "logdet", "post_mean", "post_var"):
        a, b = got[key].detach().cpu().reshape(-1), ref[key].reshape(-1)
        assert ((a - b).norm() / b.norm()).item() < 1e-6, key


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-6), (torch.float32, 1e-3)])
def test_vbagg_elbo_mid_scale_vs_oracle(dtype, tol):
    """Fused VBagg path (bag-summed cross Gram, within-bag block sums, whitened KL) on 2000 individuals in
    ragged bags vs the dense reference algorithm (oracle) - values and gradients."""
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, mlls as E, models as M
    torch.manual_seed(7)
    sizes = [int(s) for s in torch.randint(1, 120, (40,))]
    N, M_, d = sum(sizes), 48, 3
    x = torch.randn(N, d, dtype=torch.float64)
    z = torch.randn(len(sizes), dtype=torch.float64)
    Z = torch.randn(M_, d, dtype=torch.float64)
    vm = 0.2 * torch.randn(M_, dtype=torch.float64)
    cv = torch.eye(M_, dtype=torch.float64) + 0.1 * torch.randn(M_, M_, dtype=torch.float64).tril()
    raw_ls, raw_os, raw_noise = torch.tensor([0.3, -0.1, 0.5], dtype=torch.float64), torch.tensor(0.2, dtype=torch.float64), torch.tensor([-0.4], dtype=torch.float64)
    # oracle
    leaves = [t.clone().requires_grad_(True) for t in (Z, vm, cv, raw_ls, raw_os, raw_noise)]
    k = R.KernelSpec([R.Component("rbf", leaves[3], leaves[4], None)])
    mu_q, Sig = R.svgp_q(k, leaves[0], x, leaves[1], leaves[2])
    ell = R.vbagg_ell(z, mu_q, Sig, sizes, leaves[5])
    loss_ref = -R.bag_elbo(ell, R.whitened_kl(leaves[1], leaves[2]), len(sizes), len(sizes), 1.0)
    loss_ref.backward()
    # device
    km = K.ScaleKernel(K.RBFKernel(ard_num_dims=d))
    model = M.VariationalGP(inducing_points=Z.clone(), mean_module=Mn.ZeroMean(), covar_module=km)
    lik = L.VBaggGaussianLikelihood()
    model, lik = model.to(DEV).to(dtype), lik.to(DEV).to(dtype)
    with torch.no_grad():
        km.base_kernel.raw_lengthscale.copy_(raw_ls.reshape(1, -1))
        km.raw_outputscale.copy_(raw_os)
        lik.noise_covar.raw_noise.copy_(raw_noise)
        vd = model.variational_strategy._variational_distribution
        vd.variational_mean.copy_(vm)
        vd.chol_variational_covar.copy_(cv)
    model.variational_strategy.variational_params_initialized.fill_(1)
    model.train(); lik.train()
    elbo = E.BagVariationalELBO(lik, model, num_data=len(sizes), beta=1.0)
    loss = -elbo(variational_dist_f=model(x.to(DEV).to(dtype)), target=z.to(DEV).to(dtype), bags_sizes=sizes)
    loss.backward()
    rel = lambda a, b: ((a.detach().double().cpu().reshape(-1) - b.reshape(-1)).norm() / b.norm().clamp_min(1e-12)).item()
    assert rel(loss, loss_ref.detach()) < tol
    if dtype == torch.float64:
        assert rel(model.variational_strategy.inducing_points.grad, leaves[0].grad) < 1e-6
        assert rel(vd.variational_mean.grad, leaves[1].grad) < 1e-6
        assert rel(vd.chol_variational_covar.grad, leaves[2].grad) < 1e-6
        assert rel(km.base_kernel.raw_lengthscale.grad, leaves[3].grad) < 1e-6
        assert rel(km.raw_outputscale.grad, leaves[4].grad) < 1e-6
        assert rel(lik.noise_covar.raw_noise.grad, leaves[5].grad) < 1e-6


def test_row_block_sharded_kernel_matmul_equals_unsharded():
    """Panels computed rank by rank (ranks emulated in one process) and concatenated == unsharded K W."""
    from deconditional_downscaling_b200 import kernels as K
    from deconditional_downscaling_b200.parallel import row_block
    torch.manual_seed(5)
    dt = torch.float64
    x = torch.randn(1000, 3, dtype=dt, device=DEV)
    W = torch.randn(1000, 17, dtype=dt, device=DEV)
    km = K.ScaleKernel(K.RBFKernel(ard_num_dims=3)).to(DEV).double()
    with torch.no_grad():
        full = km.ops(x).matmul(x, None, W)
        parts = []
        for r in range(4):
            lo, hi = row_block(1000, r, 4)
            parts.append(km.ops(x).matmul(x[lo:hi], x, W))
        ref = R.gram(R.KernelSpec([R.Component("rbf", km.base_kernel.raw_lengthscale.detach().cpu().reshape(-1),
                                               km.raw_outputscale.detach().cpu(), None)]), x.cpu()) @ W.cpu()
    assert ((torch.cat(parts) - full).norm() / full.norm()).item() < 1e-13
    assert ((full.cpu() - ref).norm() / ref.norm()).item() < 1e-10


def test_row_block_sharded_self_product_keeps_delta_and_nugget_terms():
    """k = Scale(RBF) + Scale(Delta) (src/models/cmp.py:61-62) and a nugget: the row panels of a self-product are
    cross products for the Gram kernel, so the diagonal terms are added per panel - the result must not depend on
    the number of ranks (ranks emulated in one process)."""
    from deconditional_downscaling_b200 import kernels as K
    from deconditional_downscaling_b200.parallel import kernel_matmul_panel, row_block
    torch.manual_seed(6)
    dt = torch.float64
    x = torch.randn(700, 3, dtype=dt, device=DEV)
    W = torch.randn(700, 9, dtype=dt, device=DEV)
    km = (K.ScaleKernel(K.RBFKernel(ard_num_dims=3)) + K.ScaleKernel(K.DeltaKernel())).to(DEV).double()
    with torch.no_grad():
        kops = km.ops(x)
        full = kops.matmul(x, None, W, diag_const=0.3)
        dense = kops.gram(x, None, diag_const=0.3) @ W
        for world in (2, 3):
            parts = [kernel_matmul_panel(kops, x, None, W, *row_block(700, r, world), diag_const=0.3) for r in range(world)]
            assert ((torch.cat(parts) - full).norm() / full.norm()).item() < 1e-13
    assert ((full - dense).norm() / dense.norm()).item() < 1e-12
    delta_part = torch.nn.functional.softplus(km.kernels[1].raw_outputscale.detach()) + 0.3
    assert delta_part.item() > 0.5          # the diagonal terms are a visible part of the product


def test_gram_matmul_multi_panel_matches_single_products():
    """Above 1 GiB the Gram matrix is produced in 512 MiB row panels; (K + c I) W must not change."""
    o = ops()
    torch.manual_seed(9)
    dt = torch.float64
    n, m = 12000, 24
    x = torch.randn(n, 3, dtype=dt, device=DEV)
    W = torch.randn(n, m, dtype=dt, device=DEV)
    got = o.gram_matmul(rbf(3, dt), x, None, W, diag_const=0.37)
    ref = torch.cat([o.gemm(o.gram(rbf(3, dt), x[i:i + 3000], x), W) for i in range(0, n, 3000)]) + 0.37 * W
    assert ((got - ref).norm() / ref.norm()).item() < 1e-13


def test_evaluation_over_large_grid_is_chunk_invariant_and_matches_oracle():
    """q(f) mean / variance over 120 000 prediction points (the evaluation path of the downscaling trainer):
    chunked evaluation == one-shot evaluation, and a 3000-point slice matches the dense oracle formulas."""
    from deconditional_downscaling_b200 import evaluation, kernels as K, means as Mn, models as M
    torch.manual_seed(13)
    dt = torch.float32
    Mz, d, N = 256, 3, 120_000
    km = K.ScaleKernel(K.RBFKernel(ard_num_dims=d))
    model = M.VariationalGP(inducing_points=torch.randn(Mz, d), mean_module=Mn.ZeroMean(), covar_module=km).to(DEV).to(dt)
    vd = model.variational_strategy._variational_distribution
    with torch.no_grad():
        vd.variational_mean.copy_(0.3 * torch.randn(Mz))
        vd.chol_variational_covar.copy_(torch.eye(Mz) + 0.05 * torch.randn(Mz, Mz).tril())
    model.variational_strategy.variational_params_initialized.fill_(1)
    model.eval()
    x = torch.randn(N, d, device=DEV, dtype=dt)
    m1, v1 = evaluation.posterior_mean_variance(model, x, chunk=N)
    m2, v2 = evaluation.posterior_mean_variance(model, x, chunk=17_000)
    # FP32 models compute the wide products in single-pass TF32 (tolerance of that mode: 1e-3); the chunking
    # changes tile shapes and summation order, not the algorithm
    assert ((m1 - m2).norm() / m1.norm()).item() < 1e-3 and ((v1 - v2).norm() / v1.norm()).item() < 1e-3
    assert torch.all(v1 > 0)
    # oracle on a slice (float64, dense)
    xs = x[:3000].double().cpu()
    k = R.KernelSpec([R.Component("rbf", km.base_kernel.raw_lengthscale.detach().double().cpu().reshape(-1),
                                  km.raw_outputscale.detach().double().cpu(), None)])
    mu, Sig = R.svgp_q(k, model.variational_strategy.inducing_points.detach().double().cpu(), xs,
                       vd.variational_mean.detach().double().cpu(), vd.chol_variational_covar.detach().double().cpu())
    assert ((m1[:3000].double().cpu() - mu).norm() / mu.norm()).item() < 1e-3
    assert ((v1[:3000].double().cpu() - Sig.diagonal()).norm() / Sig.diagonal().norm()).item() < 1e-3
    # chunked NLL of the reference's metrics code on the exact-GP path
    # (well separated points: the 2000 x 2000 FP32 posterior covariance of a chunk must stay positive definite;
    # when it is not, the reference's metric is NaN by design, core/metrics.py:30-31)
    xw = 4.0 * x[:4000]
    nll = evaluation.chunked_nll(lambda xc: model(xc), xw, torch.zeros(4000, device=DEV, dtype=dt), chunk_size=2000)
    assert math.isfinite(nll)
    post = model(xw[:2000])
    cov = post.covariance_matrix.double()
    ref = -torch.distributions.MultivariateNormal(post.mean.double(), 0.5 * (cov + cov.t()), validate_args=False).log_prob(
        torch.zeros(2000, device=DEV, dtype=torch.float64)) / 2000
    nll0 = evaluation.chunked_nll(lambda xc: model(xc), xw[:2000], torch.zeros(2000, device=DEV, dtype=dt), chunk_size=2000)
    assert abs(nll0 - ref.item()) < 1e-3 * abs(ref.item())


@pytest.mark.parametrize("transa,m,n,k", [(False, 5037, 9005, 777), (True, 4993, 8200, 1031), (False, 385, 40000, 200)])
def test_cluster_multicast_gemm_ragged_shapes(transa, m, n, k):
    """Non-transposed B with more than a wave of wide tiles runs on two-CTA clusters with multicast B tiles: ragged
    M / N / K, an odd number of tile rows (the second CTA of the last pair has no tile), beta and lower-only."""
    o = ops()
    torch.manual_seed(m)
    A = torch.randn((k, m) if transa else (m, k), device=DEV)
    B = torch.randn(k, n, device=DEV)
    C0 = torch.randn(m, n, device=DEV)
    ref = (A.t() if transa else A).double() @ B.double()
    o.TF32_MODE = "fast"
    try:
        C = o.gemm(A, B, transa=transa)
        assert ((C.double() - ref).norm() / ref.norm()).item() < 2e-3
        C = o.gemm(A, B, transa=transa, alpha=-0.5, beta=1.0, C=C0.clone())
        full = C0.double() - 0.5 * ref
        assert ((C.double() - full).norm() / full.norm()).item() < 2e-3
    finally:
        o.TF32_MODE = "auto"
    X = o.gemm_x3(A, B, transa=transa)                       # fat 3xTF32 stages on clusters
    assert ((X.double() - ref).norm() / ref.norm()).item() < 1e-5
    if m <= n:
        Cl = o.gemm(A, B[:, :m].contiguous(), transa=transa, alpha=1.0, beta=1.0, C=C0[:, :m].clone(), lower=True)
        want = (C0[:, :m].double() + ref[:, :m]).tril()
        assert ((Cl.double().tril() - want).norm() / want.norm()).item() < 2e-3
