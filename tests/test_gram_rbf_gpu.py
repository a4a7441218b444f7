"""Single-component RBF fast path of the dense Gram build and its hyper-parameter adjoint (csrc/gram_rbf.cuh) against
closed-form FP64 torch expressions of the reference's kernel  os * exp(-|x - y|_l^2 / 2)  (gpytorch ScaleKernel(RBFKernel),
as built in experiments/swiss_roll/models/exact_cmp.py:35-41) and against the generic kernels (DCGP_GRAM_RBF_FAST=0)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spec(ls, os_, dtype):
    from deconditional_downscaling_b200 import ops
    d = ls.numel()
    return ops.FlatSpec(["rbf"], [list(range(d))], [ls.to(DEV, dtype)], [os_.to(DEV, dtype)])


def _dense(x1, x2, ls, os_):
    a, b = x1.double() / ls.double(), x2.double() / ls.double()
    d2 = (a[:, None, :] - b[None, :, :]).pow(2).sum(-1)
    return os_.double() * torch.exp(-0.5 * d2)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 2e-15), (torch.float32, 2e-5)])
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8])
def test_forward_every_dimension_and_ragged_layouts(dtype, tol, d):
    from deconditional_downscaling_b200 import ops
    g = torch.Generator().manual_seed(d)
    for n1, n2 in [(1, 1), (130, 257), (515, 64)]:
        x1 = torch.randn(n1, d, generator=g, dtype=torch.float64).to(dtype)
        x2 = torch.randn(n2, d, generator=g, dtype=torch.float64).to(dtype)
        ls = (0.5 + torch.rand(d, generator=g, dtype=torch.float64)).to(dtype)
        os_ = torch.tensor(1.7, dtype=dtype)
        K = ops.gram(_spec(ls, os_, dtype), x1.to(DEV), x2.to(DEV))
        ref = _dense(x1, x2, ls, os_)
        # element-wise relative error: the argument of the exponential carries the rounding of the scaled coordinates
        err = ((K.double().cpu() - ref).abs() / ref).max().item()
        assert err < tol * (1 + ref.log().abs().max().item()), (d, n1, n2, err)


def test_fp64_exponential_over_the_whole_range():
    """Arguments from 0 down to -800: <= 4 ulp wherever the result is a normal number, zero (not garbage) below."""
    from deconditional_downscaling_b200 import ops
    n = 4096
    x1 = torch.zeros(1, 1, dtype=torch.float64)
    x2 = (torch.linspace(0, 40.0, n, dtype=torch.float64)).reshape(n, 1)     # -x^2 / 2 in [-800, 0]
    ls, os_ = torch.ones(1, dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64)
    K = ops.gram(_spec(ls, os_, torch.float64), x1.to(DEV), x2.to(DEV)).cpu().reshape(-1)
    # the argument exactly as the kernel forms it: coordinate times 1 / sqrt 2 (rounded), squared (rounded once)
    dsc = x2.reshape(-1) * 0.70710678118654752
    arg = -(dsc * dsc)
    ref = torch.exp(arg)
    normal = arg > -700.0
    rel = ((K[normal] - ref[normal]).abs() / ref[normal]).max().item()
    assert rel < 4 * 2.2e-16, rel
    assert torch.all(K[~normal] >= 0) and torch.all(K[~normal] < 1e-300)
    assert torch.all(K[arg < -709.0] == 0)
    # far apart points (k leaves 32 bits) and an outputscale that is not one
    far = torch.tensor([[0.0], [1e6], [-3e8]], dtype=torch.float64)
    Kf = ops.gram(_spec(ls, torch.tensor(2.5, dtype=torch.float64), torch.float64), far.to(DEV)).cpu()
    assert (Kf.diagonal() - 2.5).abs().max().item() < 1e-15
    assert (Kf - torch.diag(Kf.diagonal())).abs().max().item() == 0.0


def test_diagonal_constant_and_symmetric_build():
    from deconditional_downscaling_b200 import ops
    for dtype, tol in [(torch.float64, 1e-14), (torch.float32, 1e-5)]:
        x = torch.randn(777, 3, dtype=torch.float64).to(dtype)
        ls, os_ = torch.tensor([0.7, 1.1, 2.0], dtype=dtype), torch.tensor(0.9, dtype=dtype)
        K = ops.gram(_spec(ls, os_, dtype), x.to(DEV), None, diag_const=0.25)
        ref = _dense(x, x, ls, os_) + 0.25 * torch.eye(777, dtype=torch.float64)
        assert ((K.double().cpu() - ref).abs().max().item()) < tol
        # exact diagonal: os + constant, whatever the coordinates
        assert (K.diagonal().double().cpu() - (os_.double() + 0.25)).abs().max().item() < (1e-15 if dtype == torch.float64 else 2e-7)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-4)])
@pytest.mark.parametrize("n1,n2,d", [(300, 200, 3), (1030, 517, 5), (64, 2048, 8), (5, 3, 1)])
def test_hyper_adjoint_against_autograd(dtype, tol, n1, n2, d):
    from deconditional_downscaling_b200 import ops
    g = torch.Generator().manual_seed(n1 + d)
    x1 = torch.randn(n1, d, generator=g, dtype=torch.float64).to(dtype)
    x2 = torch.randn(n2, d, generator=g, dtype=torch.float64).to(dtype)
    Gm = torch.randn(n1, n2, generator=g, dtype=torch.float64).to(dtype)
    ls = (0.8 + torch.rand(d, generator=g, dtype=torch.float64)).to(dtype)
    os_ = torch.tensor(1.3, dtype=dtype)
    ls64, os64 = ls.double().requires_grad_(True), os_.double().requires_grad_(True)
    (_dense(x1, x2, ls64, os64) * Gm.double()).sum().backward()
    dls, dos, _ = ops.gram_bwd(_spec(ls, os_, dtype), x1.to(DEV), x2.to(DEV), Gm.to(DEV))
    # sums of n1 * n2 signed terms: judged against the sum of magnitudes
    scale_l = (_dense(x1, x2, ls.double(), os_.double()) * Gm.double().abs()).sum().item()
    assert (dls[0, :d].double().cpu() - ls64.grad).abs().max().item() < tol * scale_l
    assert abs(dos[0].item() - os64.grad.item()) < tol * scale_l


def test_generic_kernels_agree_with_the_fast_path():
    """Same calls with DCGP_GRAM_RBF_FAST=0 (the generic additive-kernel code) in a child process."""
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from deconditional_downscaling_b200 import ops
torch.manual_seed(0)
for dt in (torch.float64, torch.float32):
    x = torch.randn(700, 5, dtype=dt, device="cuda"); G = torch.randn(700, 700, dtype=dt, device="cuda")
    sp = ops.FlatSpec(["rbf"], [list(range(5))], [torch.full((5,), 1.3, dtype=dt, device="cuda")], [torch.tensor(0.8, dtype=dt, device="cuda")])
    K = ops.gram(sp, x, None, diag_const=0.1)
    dls, dos, _ = ops.gram_bwd(sp, x, None, G)
    print(repr(K.double().sum().item()), repr(K.double().pow(2).sum().item()), " ".join(repr(v) for v in dls[0, :5].double().tolist()), repr(dos[0].item()))
''' % ROOT
    outs = []
    for flag in ("1", "0"):
        env = dict(os.environ, DCGP_GRAM_RBF_FAST=flag)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([[float(v) for v in line.split()] for line in r.stdout.strip().splitlines()])
    for row_fast, row_gen, tol in zip(outs[0], outs[1], (1e-12, 2e-4)):
        for a, b in zip(row_fast, row_gen):
            assert abs(a - b) <= tol * max(abs(b), 1.0) * 50, (row_fast, row_gen)
