"""The factorisation leaves build the inverse of their factor inside the pivot loop (csrc/chol.cu, inv_rows_*; default since
round 2 after the B200 A/B in profiles/r02_finv_ab.txt).  DCGP_LEAF_FINV=0 keeps the round-1 variant (separate block_trtri
pass) for A/B runs: this test keeps that switch honest.  The switch is read once per process, hence the subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r'''
import sys, torch
sys.path.insert(0, sys.argv[1])
from deconditional_downscaling_b200 import ops
torch.manual_seed(0)
for dt, tol in ((torch.float64, 1e-10), (torch.float32, 2e-5)):
    for n in (7, 16, 100, 128, 129, 517, 2048):
        M = torch.randn(n, n, dtype=torch.float64, device="cuda")
        A = (M @ M.t() / n + torch.eye(n, dtype=torch.float64, device="cuda")).to(dt)
        L = A.clone()
        info = ops.potrf_(L)
        assert int(info.item()) == 0, (dt, n, int(info.item()))
        L = L.tril()
        ref = torch.linalg.cholesky(A.double())
        err = ((L.double() - ref).norm() / ref.norm()).item()
        assert err < tol, ("factor", dt, n, err)
        B = torch.randn(n, 37, dtype=dt, device="cuda")
        X = B.clone()
        ops.potrs_(L, X)                       # uses the inverted diagonal blocks the leaves produced
        res = ((A.double() @ X.double() - B.double()).norm() / B.double().norm()).item()
        assert res < 50 * tol, ("solve", dt, n, res)
        W = ops.trtri(L)
        idn = ((W.double() @ L.double() - torch.eye(n, dtype=torch.float64, device="cuda")).norm() / n ** 0.5).item()
        assert idn < 50 * tol, ("trtri", dt, n, idn)
print("ok")
'''


def test_factorisation_with_legacy_leaves_matches_lapack():
    """DCGP_LEAF_FINV=0: factors against LAPACK, solves and inverses through the blocks the leaves produced (the default
    mode is what every other GPU test exercises)."""
    r = subprocess.run([sys.executable, "-c", CHECK, ROOT], env={**os.environ, "DCGP_LEAF_FINV": "0"},
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-3000:]


CHECK_LA2 = r'''
import sys, torch
sys.path.insert(0, sys.argv[1])
from deconditional_downscaling_b200 import ops
torch.manual_seed(1)
for dt, tol in ((torch.float64, 1e-11), (torch.float32, 1e-4)):
    for n, m in ((2048, 64), (2561, 100), (3583, 200), (5000, 520)):
        M = torch.randn(n, n, dtype=torch.float64, device="cuda")
        A = M @ M.t() / n + torch.eye(n, dtype=torch.float64, device="cuda")
        L = torch.linalg.cholesky(A).to(dt).contiguous()
        plan = ops.solve_plan(L)
        B = torch.randn(n, m, dtype=dt, device="cuda")
        Ld, Bd = L.double(), B.double()
        rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
        assert rel(ops.trsm_(L, B.clone(), plan=plan), torch.linalg.solve_triangular(Ld, Bd, upper=False)) < tol, (dt, n, m, "N")
        assert rel(ops.trsm_(L, B.clone(), trans=True, plan=plan), torch.linalg.solve_triangular(Ld.t(), Bd, upper=True)) < tol, (dt, n, m, "T")
        assert rel(ops.potrs_(L, B.clone(), plan=plan), torch.cholesky_solve(Bd, Ld)) < 10 * tol, (dt, n, m, "potrs")
        assert rel(ops.potrs_(L, B.clone()), torch.cholesky_solve(Bd, Ld)) < 10 * tol, (dt, n, m, "unplanned")
print("ok")
'''


def test_two_step_lookahead_sweeps_match_lapack():
    """DCGP_TRSM_LA=2 (opt-in, csrc/chol.cu ltrsm512_la2: one dependent GEMM per 512-block on the chain lane, products
    with the inverted neighbour blocks prepared in the plan): exactly four blocks, ragged last blocks of 1 and 511 rows,
    many blocks; planned and unplanned; forward, transposed and both sweeps."""
    r = subprocess.run([sys.executable, "-c", CHECK_LA2, ROOT], env={**os.environ, "DCGP_TRSM_LA": "2"},
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-3000:]


CHECK_HELPER = r'''
import sys, torch
sys.path.insert(0, sys.argv[1])
from deconditional_downscaling_b200 import ops
torch.manual_seed(2)
for n in (6144, 6271, 8192):
    x = torch.randn(n, 3, device="cuda")
    spec = ops.FlatSpec(["rbf"], [[0, 1, 2]], [torch.ones(3, device="cuda")], [torch.ones(1, device="cuda")])
    A = ops.gram(spec, x, None, diag_const=0.8)
    L = A.clone()
    info = ops.potrf_(L)
    assert int(info.item()) == 0, n
    L = L.tril().double()
    v = torch.randn(n, 4, dtype=torch.float64, device="cuda")
    ref = A.double() @ v
    err = ((L @ (L.t() @ v) - ref).norm() / ref.norm()).item()
    assert err < 1e-5, (n, err)
    Lr = torch.linalg.cholesky(A.double())
    assert ((L - Lr).norm() / Lr.norm()).item() < 1e-4, n
print("ok")
'''


def test_helper_lane_factorisation_matches_lapack():
    """DCGP_POTRF_H=14 (opt-in, csrc/chol.cu chol_la_fused: panel solves and next-panel strips on a helper lane with its own
    share of the SMs; B200: 8192^2 3.51 -> 3.31 ms stand-alone, neutral inside the ELBO step): full and ragged sizes from
    the 48-block threshold up."""
    r = subprocess.run([sys.executable, "-c", CHECK_HELPER, ROOT], env={**os.environ, "DCGP_POTRF_H": "14"},
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-3000:]
