"""GPU parity tests of the C-ABI kernels (through the ctypes launchers in
deconditional_downscaling_b200.ops) against the CPU oracle (oracle/restate.py)
and the golden vectors produced by the reference's own src/ code.

Tolerances: FP64 path 1e-6 relative (north_star), stated per assert; the
FP32/TF32 path 1e-3 relative against the same FP64 oracle values.
"""
import math

import pytest
import torch
from torch.nn.functional import softplus

import golden_util as G
from oracle import restate as R

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def ops():
    from deconditional_downscaling_b200 import ops as o
    return o


def dev(t, dtype):
    return t.to(device=DEV, dtype=dtype)


def flat(spec, dtype):
    o = ops()
    return o.FlatSpec([c.kind for c in spec.components],
                      [list(c.dims) if c.dims is not None else list(range(c.raw_lengthscale.numel())) for c in spec.components],
                      [dev(c.lengthscale.detach(), dtype) for c in spec.components],
                      [dev(c.outputscale.detach(), dtype) for c in spec.components])


def comp(kind, ls, os_, dims=None):
    return R.Component(kind, torch.as_tensor(ls, dtype=torch.float64), torch.as_tensor(os_, dtype=torch.float64), dims)


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


TOLS = {torch.float64: 1e-6, torch.float32: 1e-3}


# ------------------------------------------------------------------ Gram forward
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gram_matches_reference_golden(dtype):
    g = G.load("gram")
    rbf = R.KernelSpec([comp("rbf", g["rbf_raw_lengthscale"], g["rbf_raw_outputscale"])])
    mat = R.KernelSpec([comp("matern32", g["mat_raw_lengthscale"], g["mat_raw_outputscale"], dims=[0, 1])])
    add = R.KernelSpec(mat.components + [comp("rbf", g["add_rbf_raw_lengthscale"], g["add_rbf_raw_outputscale"], dims=[2])])
    for name, s in (("rbf", rbf), ("mat", mat), ("add", add)):
        K12 = ops().gram(flat(s, dtype), dev(g["x1"], dtype), dev(g["x2"], dtype))
        K11 = ops().gram(flat(s, dtype), dev(g["x1"], dtype))
        tol = 1e-10 if dtype == torch.float64 else 2e-5   # fp32 here is plain FFMA + __expf, not TF32
        assert relerr(K12, g[name + "_x1x2"]) < tol
        assert relerr(K11, g[name + "_x1x1"]) < tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n1,n2,d", [(1, 1, 1), (63, 65, 3), (257, 130, 5), (1000, 777, 8)])
def test_gram_ragged_shapes_vs_oracle(dtype, n1, n2, d):
    torch.manual_seed(n1 * 7 + n2)
    x1, x2 = torch.randn(n1, d, dtype=torch.float64), torch.randn(n2, d, dtype=torch.float64)
    s = R.KernelSpec([comp("matern32", torch.randn(d) * 0.3, 0.1)]) if d % 2 else R.KernelSpec([comp("rbf", torch.randn(d) * 0.3, -0.2)])
    ref = R.gram(s, x1, x2)
    K = ops().gram(flat(s, dtype), dev(x1, dtype), dev(x2, dtype))
    assert relerr(K, ref) < (1e-10 if dtype == torch.float64 else 3e-5)


def test_gram_diag_add_and_exact_diagonal():
    torch.manual_seed(0)
    x = torch.randn(300, 3, dtype=torch.float64)
    s = R.KernelSpec([comp("rbf", [0.1, 0.2, 0.3], 0.4)])
    extra = torch.tensor([0.25], dtype=torch.float64, device=DEV)
    K = ops().gram(flat(s, torch.float64), dev(x, torch.float64), None, diag_dev=extra, diag_const=1.5)
    ref = R.gram(s, x) + (1.5 + 0.25) * torch.eye(300, dtype=torch.float64)
    assert relerr(K, ref) < 1e-12
    # difference form => the diagonal is exactly outputscale + diag (the reference zero-fills it, Appendix A)
    os_ = softplus(torch.tensor(0.4, dtype=torch.float64)).item()
    assert torch.equal(K.diagonal().cpu(), torch.full((300,), os_ + 1.75, dtype=torch.float64))


# ------------------------------------------------------------------ Gram backward
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("case", ["rbf3", "additive", "same"])
def test_gram_bwd_vs_oracle_autograd(dtype, case):
    torch.manual_seed(5)
    n1, n2 = 150, 97
    if case == "additive":
        d = 5
        raw = [torch.randn(2, dtype=torch.float64).requires_grad_(), torch.randn(3, dtype=torch.float64).requires_grad_()]
        ros = [torch.tensor(0.3, dtype=torch.float64, requires_grad=True), torch.tensor(-0.2, dtype=torch.float64, requires_grad=True)]
        s = R.KernelSpec([R.Component("matern32", raw[0], ros[0], [0, 1]), R.Component("rbf", raw[1], ros[1], [2, 3, 4])])
    else:
        d = 3
        raw = [torch.randn(3, dtype=torch.float64).requires_grad_()]
        ros = [torch.tensor(0.3, dtype=torch.float64, requires_grad=True)]
        s = R.KernelSpec([R.Component("rbf", raw[0], ros[0], None)])
    x1 = torch.randn(n1, d, dtype=torch.float64, requires_grad=True)
    same = case == "same"
    x2 = None if same else torch.randn(n2, d, dtype=torch.float64)
    Gup = torch.randn(n1, n1 if same else n2, dtype=torch.float64)
    # oracle: gradient w.r.t. the *transformed* hypers (lengthscale, outputscale) and x1
    ls = [softplus(r).detach().requires_grad_() for r in raw]
    os_ = [softplus(r).detach().requires_grad_() for r in ros]

    def k_of(a, b):
        out = 0
        for c, l, o in zip(s.components, ls, os_):
            aa, bb = R._select(a, c.dims), R._select(b, c.dims)
            dist2 = ((aa.unsqueeze(1) - bb.unsqueeze(0)) / l).pow(2).sum(-1)
            if c.kind == "rbf":
                out = out + o * torch.exp(-0.5 * dist2)
            else:
                r = dist2.clamp_min(1e-30).sqrt()
                out = out + o * (1 + math.sqrt(3) * r) * torch.exp(-math.sqrt(3) * r)
        return out

    K = k_of(x1, x1 if same else x2)
    (K * Gup).sum().backward()
    fs = flat(s, dtype)
    dls, dos, dx1 = ops().gram_bwd(fs, dev(x1.detach(), dtype), None if same else dev(x2, dtype), dev(Gup, dtype), need_dx1=True)
    tol = 1e-9 if dtype == torch.float64 else 2e-4
    for c in range(len(ls)):
        nd = ls[c].numel()
        assert relerr(dls[c, :nd], ls[c].grad) < tol
        assert relerr(dos[c], os_[c].grad) < tol
    assert relerr(dx1, x1.grad) < tol


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (130, 70, 33), (257, 300, 515), (64, 1000, 128)])
def test_gemm_all_layouts(dtype, ta, tb, m, n, k):
    torch.manual_seed(m + n + k)
    A = torch.randn((k, m) if ta else (m, k), dtype=torch.float64)
    B = torch.randn((n, k) if tb else (k, n), dtype=torch.float64)
    C0 = torch.randn(m, n, dtype=torch.float64)
    ref = 0.7 * (A.t() if ta else A) @ (B.t() if tb else B) - 0.3 * C0
    C = dev(C0, dtype).clone()
    ops().gemm(dev(A, dtype), dev(B, dtype), transa=ta, transb=tb, alpha=0.7, beta=-0.3, C=C)
    assert relerr(C, ref) < (1e-13 if dtype == torch.float64 else 2e-3)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gemm_split_k_and_lower(dtype):
    torch.manual_seed(1)
    A = torch.randn(200, 4096, dtype=torch.float64)
    ref = A @ A.t()
    C = ops().gemm(dev(A, dtype), dev(A, dtype), transb=True)     # 2x2 tiles, deep K => split-K path
    assert relerr(C, ref) < (1e-13 if dtype == torch.float64 else 2e-3)
    P = torch.randn(700, 128, dtype=torch.float64)
    S0 = torch.randn(700, 700, dtype=torch.float64)
    S = dev(S0, dtype).clone()
    ops().gemm(dev(P, dtype), dev(P, dtype), transb=True, alpha=-1.0, beta=1.0, C=S, lower=True)
    ref = S0 - P @ P.t()
    assert relerr(S.tril(), ref.tril()) < (1e-13 if dtype == torch.float64 else 2e-3)
    # tiles strictly above the diagonal are untouched
    assert torch.equal(S[:128, 128:].cpu().double(), S0[:128, 128:].to(dtype).double())


# ------------------------------------------------------------------ Cholesky / solves
def spd(n, dtype=torch.float64, seed=0):
    torch.manual_seed(seed)
    x = torch.randn(n, 3, dtype=torch.float64)
    s = R.KernelSpec([comp("rbf", [0.5, 0.5, 0.5], 0.5)])
    return R.gram(s, x) + 0.05 * n ** 0.5 * torch.eye(n, dtype=torch.float64)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n", [1, 7, 128, 129, 300, 1000])
def test_potrf_vs_lapack(dtype, n):
    A = spd(n)
    ref = torch.linalg.cholesky(A)
    L = dev(A, dtype).clone()
    info = ops().potrf_(L)
    assert info.item() == 0
    L = L.tril()
    assert relerr(L, ref) < (1e-12 if dtype == torch.float64 else 2e-3)
    assert relerr(L.double() @ L.double().t(), A) < (1e-13 if dtype == torch.float64 else 2e-3)
    # logdet
    ld = ops().logdet_chol(L)
    assert abs(ld.item() - torch.logdet(A).item()) < (1e-9 if dtype == torch.float64 else 1e-3) * max(1.0, abs(torch.logdet(A).item()))


def test_potrf_reports_first_bad_pivot():
    A = spd(300)
    A[200, 200] = -1.0
    L = dev(A, torch.float64).clone()
    info = ops().potrf_(L)
    assert info.item() == 201


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n,m", [(5, 1), (128, 40), (300, 129), (1000, 64), (1500, 70), (2048, 64), (2300, 200),
                                 (1024, 256), (1536, 512)])
def test_trsm_and_potrs_vs_lapack(dtype, n, m):
    A = spd(n, seed=n)
    Lref = torch.linalg.cholesky(A)
    torch.manual_seed(m)
    B = torch.randn(n, m, dtype=torch.float64)
    L = dev(Lref, dtype)
    # FP32 solves run error-compensated (3xTF32, mma.sync and tcgen05 kernels): FP32-level accuracy
    tol = 1e-11 if dtype == torch.float64 else 1e-4
    X0 = ops().trsm_(L, dev(B, dtype).clone(), trans=False)
    assert relerr(X0, torch.linalg.solve_triangular(Lref, B, upper=False)) < tol
    X1 = ops().trsm_(L, dev(B, dtype).clone(), trans=True)
    assert relerr(X1, torch.linalg.solve_triangular(Lref.t(), B, upper=True)) < tol
    X2 = ops().potrs_(L, dev(B, dtype).clone())
    assert relerr(X2, torch.cholesky_solve(B, Lref)) < tol
    # size-independent property: residual of the solve
    assert relerr(dev(A, torch.float64) @ X2.double(), B) < (1e-11 if dtype == torch.float64 else 3e-3)


# ------------------------------------------------------------------ bag kernels
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_bag_kernels_vs_oracle(dtype):
    torch.manual_seed(9)
    sizes = [5, 1, 300, 17, 64, 2]
    N, M, d = sum(sizes), 140, 3
    offs = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    x, z = torch.randn(N, d, dtype=torch.float64), torch.randn(M, d, dtype=torch.float64)
    raw = torch.randn(3, dtype=torch.float64)
    s = R.KernelSpec([comp("rbf", raw, 0.2)])
    Kzx, Kxx = R.gram(s, z, x), R.gram(s, x)
    E = torch.zeros(N, len(sizes), dtype=torch.float64)
    for a in range(len(sizes)):
        E[offs[a]:offs[a + 1], a] = 1
    fs = flat(s, dtype)
    tol = 1e-10 if dtype == torch.float64 else 5e-5
    out = ops().gram_bagsum(fs, dev(z, dtype), dev(x, dtype), offs.to(DEV))
    assert relerr(out, (Kzx @ E).t()) < tol
    bs = ops().gram_bag_blocksum(fs, dev(x, dtype), offs.to(DEV))
    assert relerr(bs, (E.t() @ Kxx @ E).diagonal()) < tol
    X = torch.randn(N, 37, dtype=torch.float64)
    assert relerr(ops().bag_sum(dev(X, dtype), offs.to(DEV)), E.t() @ X) < tol
    # backward of the bag-summed cross Gram == dense gram_bwd with G expanded over the bag
    Gup = torch.randn(len(sizes), M, dtype=torch.float64)
    dls, dos, dz = ops().gram_bagsum_bwd(fs, dev(z, dtype), dev(x, dtype), offs.to(DEV), dev(Gup, dtype))
    dls2, dos2, dz2 = ops().gram_bwd(fs, dev(z, dtype), dev(x, dtype), dev((E @ Gup).t().contiguous(), dtype), need_dx1=True)
    t2 = 1e-9 if dtype == torch.float64 else 2e-4
    assert relerr(dls, dls2) < t2 and relerr(dos, dos2) < t2 and relerr(dz, dz2) < t2
    g = torch.randn(len(sizes), dtype=torch.float64)
    dls3, dos3 = ops().gram_bag_blocksum_bwd(fs, dev(x, dtype), offs.to(DEV), dev(g, dtype))
    Gd = (E * g) @ E.t()
    dls4, dos4, _ = ops().gram_bwd(fs, dev(x, dtype), None, dev(Gd, dtype))
    assert relerr(dls3, dls4) < t2 and relerr(dos3, dos4) < t2


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_whitened_kl_vs_oracle(dtype):
    torch.manual_seed(3)
    M = 200
    m = torch.randn(M, dtype=torch.float64, requires_grad=True)
    raw = (torch.eye(M, dtype=torch.float64) + 0.1 * torch.randn(M, M, dtype=torch.float64)).requires_grad_()
    kl = R.whitened_kl(m, raw)
    kl.backward()
    out, dm, dLs = ops().whitened_kl(dev(m.detach(), dtype), dev(raw.detach(), dtype), need_grad=True)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    assert abs(out.item() - kl.item()) < tol * abs(kl.item())
    assert relerr(dm, m.grad) < tol and relerr(dLs, raw.grad) < tol


# ------------------------------------------------------------------ tcgen05 / TMA / TMEM TF32 GEMM
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("m,n,k", [(1024, 2048, 512), (1000, 1500, 328), (2052, 1028, 96)])
def test_tcgen05_gemm_all_layouts(ta, tb, m, n, k):
    """Shapes large enough to be routed to the tcgen05 kernel (gemm_tc.cu): every operand major,
    ragged M/N/K (TMA zero fill + guarded epilogue), alpha/beta epilogue."""
    o = ops()
    torch.manual_seed(m + n + k)
    A = torch.randn((k, m) if ta else (m, k), dtype=torch.float64)
    B = torch.randn((n, k) if tb else (k, n), dtype=torch.float64)
    C0 = torch.randn(m, n, dtype=torch.float64)
    ref = 0.7 * (A.t() if ta else A) @ (B.t() if tb else B) - 0.3 * C0
    Ad, Bd = dev(A, torch.float32), dev(B, torch.float32)
    o.TF32_MODE = "fast"
    try:
        C = dev(C0, torch.float32).clone()
        o.gemm(Ad, Bd, transa=ta, transb=tb, alpha=0.7, beta=-0.3, C=C)
        o.USE_TCGEN05 = False
        C2 = dev(C0, torch.float32).clone()
        o.gemm(Ad, Bd, transa=ta, transb=tb, alpha=0.7, beta=-0.3, C=C2)
    finally:
        o.USE_TCGEN05 = True
        o.TF32_MODE = "auto"
    assert relerr(C, ref) < 2e-3          # TF32 operands, FP32 accumulation
    assert relerr(C, C2) < 1e-3           # tcgen05 kernel agrees with the mma.sync kernel


def test_tcgen05_syrk_lower_tiles():
    o = ops()
    torch.manual_seed(2)
    P = torch.randn(3000, 512, dtype=torch.float64)
    S0 = torch.randn(3000, 3000, dtype=torch.float64)
    o.TF32_MODE = "fast"
    try:
        S = dev(S0, torch.float32).clone()
        o.gemm(dev(P, torch.float32), dev(P, torch.float32), transb=True, alpha=-1.0, beta=1.0, C=S, lower=True)
    finally:
        o.TF32_MODE = "auto"
    ref = S0 - P @ P.t()
    assert relerr(S.tril(), ref.tril()) < 2e-3
    # tiles strictly above the diagonal (128-row x 256-col tiles) are untouched
    assert torch.equal(S[:128, 256:].cpu().double(), S0[:128, 256:].float().double())


@pytest.mark.parametrize("dtype,n,tol", [(torch.float64, 128, 1e-12), (torch.float64, 700, 1e-11), (torch.float64, 1024, 1e-11),
                                         (torch.float32, 96, 1e-5), (torch.float32, 1024, 1e-4), (torch.float32, 1500, 1e-4)])
def test_trtri_matches_triangular_solve(dtype, n, tol):
    """dcgp_trtri: W = L^{-1}, lower, zeros above; ragged sizes go through the partial last pair."""
    from deconditional_downscaling_b200 import ops
    torch.manual_seed(n)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    A = A @ A.t() / n + torch.eye(n, dtype=torch.float64, device="cuda")
    L = torch.linalg.cholesky(A).to(dtype).contiguous()
    W = ops.trtri(L)
    ref = torch.linalg.solve_triangular(L.double(), torch.eye(n, dtype=torch.float64, device="cuda"), upper=False)
    assert torch.equal(W.triu(1), torch.zeros_like(W))
    assert ((W.double() - ref).norm() / ref.norm()).item() < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_gaussian_terms_values_and_gradients(dtype, tol):
    """Fused (log|C|, e^T C^-1 e, tr(C^-1 G)) against torch's Cholesky + solves in float64."""
    from deconditional_downscaling_b200 import functional as F
    torch.manual_seed(3)
    n = 300
    M = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C64 = (M @ M.t() / n + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")).requires_grad_(True)
    e64 = torch.randn(n, dtype=torch.float64, device="cuda", requires_grad=True)
    Gm = torch.randn(n, n, dtype=torch.float64, device="cuda")
    G64 = (Gm @ Gm.t() / n).requires_grad_(True)
    Lr = torch.linalg.cholesky(C64)
    ref = torch.stack([2 * Lr.diagonal().log().sum(), (e64 * torch.cholesky_solve(e64[:, None], Lr)[:, 0]).sum(),
                       torch.cholesky_solve(G64, Lr).diagonal().sum()])
    wts = torch.tensor([0.7, -1.3, 0.4], dtype=torch.float64, device="cuda")
    (ref * wts).sum().backward()
    C, e, G = (t.detach().to(dtype).requires_grad_(True) for t in (C64, e64, G64))
    got = torch.stack(F.gaussian_terms(C, e, G))
    (got * wts.to(dtype)).sum().backward()
    rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
    assert rel(got, ref.detach()) < tol
    assert rel(e.grad, e64.grad) < tol and rel(G.grad, G64.grad) < tol
    sym = lambda g: 0.5 * (g + g.t())
    assert rel(sym(C.grad), sym(C64.grad)) < tol


@pytest.mark.parametrize("dtype,n,m,tol", [(torch.float32, 2300, 700, 1e-4), (torch.float32, 4096, 256, 1e-4),
                                           (torch.float64, 2048, 128, 1e-11), (torch.float32, 600, 64, 1e-4),
                                           # few / oddly laid out right-hand sides: 512-block sweeps from 16 columns, FP32
                                           # through a padded copy (rows must be 16-byte aligned for TMA)
                                           (torch.float32, 2300, 50, 1e-4), (torch.float32, 3000, 17, 1e-4),
                                           (torch.float64, 2300, 33, 1e-11), (torch.float32, 2048, 7, 1e-4),
                                           # two-step look-ahead sweeps (>= 4 blocks of 512): exactly 4 blocks, a last
                                           # block of one row, of 511 rows, many blocks
                                           (torch.float32, 2561, 100, 1e-4), (torch.float64, 3583, 200, 1e-11),
                                           (torch.float32, 5000, 520, 1e-4), (torch.float64, 4096, 64, 1e-11)])
def test_planned_solves_match_reference(dtype, n, m, tol):
    """dcgp_solve_plan + dcgp_trsm_planned (trans 0 / 1 / 2) against float64 triangular solves; a plan
    is reusable across calls and small factors (no plan) fall back to dcgp_trsm."""
    from deconditional_downscaling_b200 import ops
    torch.manual_seed(n + m)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    A = A @ A.t() / n + torch.eye(n, dtype=torch.float64, device="cuda")
    L = torch.linalg.cholesky(A).to(dtype).contiguous()
    plan = ops.solve_plan(L)
    assert (plan is None) == (n < 1024)
    B = torch.randn(n, m, dtype=dtype, device="cuda")
    Ld, Bd = L.double(), B.double()
    rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
    for _ in range(2):
        assert rel(ops.trsm_(L, B.clone(), plan=plan), torch.linalg.solve_triangular(Ld, Bd, upper=False)) < tol
        assert rel(ops.trsm_(L, B.clone(), trans=True, plan=plan), torch.linalg.solve_triangular(Ld.t(), Bd, upper=True)) < tol
        assert rel(ops.potrs_(L, B.clone(), plan=plan), torch.cholesky_solve(Bd, Ld)) < 10 * tol


def test_split_matmul_whitening_matches_float64_solve():
    """FP32 whitening L_Z^{-1} K through the FP64 inverse factor carried as an FP32 hi/lo pair (3xTF32
    tensor-core products) against the float64 triangular solve, values and both gradients."""
    from deconditional_downscaling_b200 import functional as F
    torch.manual_seed(11)
    M, N = 512, 2048
    Z = torch.randn(M, 3, dtype=torch.float64, device="cuda")
    Kzz = (torch.exp(-0.5 * torch.cdist(Z, Z) ** 2) + 1e-3 * torch.eye(M, dtype=torch.float64, device="cuda")).requires_grad_(True)
    K64 = torch.randn(M, N, dtype=torch.float64, device="cuda").requires_grad_(True)
    G = torch.randn(M, N, dtype=torch.float64, device="cuda")
    ref = torch.linalg.solve_triangular(torch.linalg.cholesky(Kzz), K64, upper=False)
    (ref * G).sum().backward()
    A2 = Kzz.detach().clone().requires_grad_(True)
    K32 = K64.detach().float().requires_grad_(True)
    got = F.split_matmul(F.chol_inverse(A2), K32)
    (got * G.float()).sum().backward()
    rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
    assert got.dtype == torch.float32
    assert rel(got, ref.detach()) < 5e-6
    assert rel(K32.grad, K64.grad) < 5e-6
    sym = lambda g: 0.5 * (g + g.t())
    assert rel(sym(A2.grad), sym(Kzz.grad)) < 1e-4


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-3)])
def test_gram_matmul_equals_gram_then_gemm(dtype, tol):
    """dcgp_gram_matmul (row panels through the workspace) == materialised Gram @ W, with and without the
    nugget, for same and cross inputs, ragged sizes."""
    from deconditional_downscaling_b200 import ops
    from deconditional_downscaling_b200.exact import rbf_spec
    torch.manual_seed(2)
    spec = rbf_spec(3, dtype, "cuda")
    x = torch.randn(1111, 3, dtype=dtype, device="cuda")
    y = torch.randn(777, 3, dtype=dtype, device="cuda")
    W1 = torch.randn(1111, 65, dtype=dtype, device="cuda")
    W2 = torch.randn(777, 200, dtype=dtype, device="cuda")
    dd = torch.tensor([0.3], dtype=dtype, device="cuda")
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    ref = ops.gram(spec, x, None, diag_dev=dd, diag_const=0.2).double() @ W1.double()
    assert rel(ops.gram_matmul(spec, x, None, W1, diag_dev=dd, diag_const=0.2), ref) < tol
    ref = ops.gram(spec, x, y).double() @ W2.double()
    assert rel(ops.gram_matmul(spec, x, y, W2), ref) < tol


def test_syrk_lower_triangle():
    from deconditional_downscaling_b200 import ops
    torch.manual_seed(4)
    A = torch.randn(300, 500, dtype=torch.float64, device="cuda")
    C0 = torch.randn(300, 300, dtype=torch.float64, device="cuda")
    C = ops.syrk(A, alpha=-1.0, beta=1.0, C=C0.clone())
    ref = C0 - A @ A.t()
    assert torch.allclose(C.tril(), ref.tril(), rtol=1e-12, atol=1e-10)
    Ct = ops.syrk(A, trans=True)
    assert torch.allclose(Ct.tril(), (A.t() @ A).tril(), rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_adam_step_matches_torch_adam_and_honours_the_veto(dtype):
    """dcgp_adam_step = torch.optim.Adam (defaults) on flat buffers; a non-zero veto word leaves parameters, moments and
    the step counter untouched; grad_scale averages a summed gradient; zero_grads clears the buffer."""
    from deconditional_downscaling_b200 import ops
    torch.manual_seed(3)
    n = 70001
    p0 = torch.randn(n, dtype=dtype, device=DEV)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.01)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    step = torch.zeros(1, dtype=torch.float64, device=DEV)
    veto = torch.zeros(3, dtype=torch.int32, device=DEV)
    for it in range(5):
        g = torch.randn(n, dtype=dtype, device=DEV) * (10.0 ** (it - 2))
        ref.grad = g.clone()
        opt.step()
        g2 = 2.0 * g                                     # as after a SUM all-reduce over two ranks
        ops.adam_step(p, g2, m, v, step, 0.01, 0.9, 0.999, 1e-8, grad_scale=0.5, veto=veto, zero_grads=True)
        assert g2.abs().max().item() == 0
        tol = 1e-12 if dtype == torch.float64 else 2e-6
        assert ((p - ref.detach()).abs().max() / ref.detach().abs().max()).item() < tol
    assert step.item() == 5
    before = (p.clone(), m.clone(), v.clone())
    veto[1] = 7
    g = torch.randn(n, dtype=dtype, device=DEV)
    ops.adam_step(p, g, m, v, step, 0.01, 0.9, 0.999, 1e-8, veto=veto, zero_grads=True)
    assert step.item() == 5 and torch.equal(p, before[0]) and torch.equal(m, before[1]) and torch.equal(v, before[2])
    assert g.abs().max().item() > 0                      # a vetoed update does not clear the gradient either


@pytest.mark.parametrize("n1,n2,m,same,kind", [(4096, 4096, 1024, True, "rbf5"), (4096, 4096, 512, True, "rbf5"), (5000, 3000, 328, False, "rbf3"),
                                               (8192, 8192, 100, True, "add"), (4100, 2100, 77, False, "mat"),
                                               (2048, 2048, 32, True, "rbf3")])
def test_fused_gram_matmul_tcgen05_matches_dense_product(n1, n2, m, same, kind):
    """csrc/gram_mm.cu: (k(x1, x2) + c I) W with the Gram tiles generated in shared memory and fed to tcgen05.mma - against
    the dense FP64 product (single-pass TF32 with operands rounded to nearest: 1e-3 normwise, measured ~2e-4) and against
    the panel path (Gram rows into a workspace + GEMM), for ragged sizes, W with unaligned rows, additive / Matern
    kernels with active dims and the diagonal terms of a self-product."""
    from deconditional_downscaling_b200 import ops
    torch.manual_seed(n1 + m)
    d = {"rbf5": 5, "rbf3": 3, "add": 5, "mat": 3}[kind]
    dev = lambda t: t.to("cuda")
    if kind == "add":
        spec64 = ops.FlatSpec(["matern32", "rbf"], [[0, 1], [2, 3, 4]], [dev(torch.tensor([0.9, 1.3], dtype=torch.float64)),
                              dev(torch.tensor([1.1, 0.7, 1.5], dtype=torch.float64))],
                              [dev(torch.tensor([0.6], dtype=torch.float64)), dev(torch.tensor([0.8], dtype=torch.float64))])
    else:
        spec64 = ops.FlatSpec(["matern32" if kind == "mat" else "rbf"], [list(range(d))],
                              [dev(0.8 + 0.5 * torch.rand(d, dtype=torch.float64))], [dev(torch.tensor([0.7], dtype=torch.float64))])
    spec32 = spec64.cast(torch.float32)
    x1 = torch.randn(n1, d, dtype=torch.float64, device="cuda")
    x2 = x1 if same else torch.randn(n2, d, dtype=torch.float64, device="cuda")
    W = torch.randn(n2, m, dtype=torch.float64, device="cuda")
    diag = 0.37 if same else 0.0
    ref = ops.gram(spec64, x1, None if same else x2, diag_const=diag) @ W
    x1f, x2f, Wf = x1.float(), (None if same else x2.float()), W.float()
    assert ops.FUSED_GRAM_MATMUL
    from deconditional_downscaling_b200 import _lib as L_
    n0 = L_.load().dcgp_launch_count()
    got = ops.gram_matmul(spec32, x1f, x2f, Wf, diag_const=diag)
    launches = L_.load().dcgp_launch_count() - n0
    if m <= 512:
        assert launches <= 2, launches            # the fused kernel (+ the one-off TF32 rounding of W): no panels, no GEMM
    else:
        assert launches >= 3                      # wider products stay on panel generation + GEMM (measured faster)
    ops.FUSED_GRAM_MATMUL = False
    try:
        panel = ops.gram_matmul(spec32, x1f, x2f, Wf, diag_const=diag)
    finally:
        ops.FUSED_GRAM_MATMUL = True
    rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
    assert rel(got, ref) < 1e-3, rel(got, ref)
    assert rel(got, panel.double()) < 1e-3
    assert torch.isfinite(got).all()


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("n,d,D", [(1, 2, 8), (257, 3, 100), (1200, 2, 1000), (5000, 5, 64)])
def test_rff_features_and_adjoint_through_the_c_abi(dtype, tol, n, d, D):
    """dcgp_rff_features / dcgp_rff_features_bwd directly (src/kernels/rff_kernel.py:82-89: [cos(x W / l), sin(x W / l)]):
    values against torch in FP64, dx and dl against autograd of the same expression; ragged sizes, one row, D = 1000 as
    in experiments/downscaling/config/variational_cmp_indiv_noise.yaml."""
    from deconditional_downscaling_b200 import ops
    torch.manual_seed(n + D)
    x = torch.randn(n, d, dtype=torch.float64, device="cuda")
    W = torch.randn(d, D, dtype=torch.float64, device="cuda")
    ls = (0.5 + torch.rand(d, dtype=torch.float64, device="cuda"))
    xr, lr = x.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    proj = xr @ (W / lr.reshape(-1, 1))
    ref = torch.cat([torch.cos(proj), torch.sin(proj)], -1)
    got = ops.rff_features(x.to(dtype), W.to(dtype), ls.to(dtype))
    assert got.shape == (n, 2 * D)
    assert ((got.double() - ref.detach()).abs().max()).item() < tol * 10
    g = torch.randn(n, 2 * D, dtype=torch.float64, device="cuda")
    ref.backward(g)
    dx, dls = ops.rff_features_bwd(x.to(dtype), W.to(dtype), ls.to(dtype), g.to(dtype).contiguous(), need_dx=True)
    rel = lambda a, b: ((a.double() - b).norm() / b.norm().clamp_min(1e-300)).item()
    assert rel(dx, xr.grad) < tol * 50, rel(dx, xr.grad)
    assert rel(dls.reshape(-1), lr.grad) < tol * 50, rel(dls.reshape(-1), lr.grad)
