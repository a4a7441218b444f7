"""Small end-to-end exercise of the tcgen05 / look-ahead paths for compute-sanitizer."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
torch.manual_seed(0)
x = torch.randn(1100, 3, device="cuda")
A = ops.gram(rbf_spec(3, torch.float32, "cuda"), x, None, diag_const=0.5)
L = A.clone(); info = ops.potrf_(L); L = L.tril()
B = torch.randn(1100, 200, device="cuda")
X = ops.potrs_(L, B.clone(), plan=ops.solve_plan(L))
print("potrf info", info.item(), "resid", ((A @ X - B).norm() / B.norm()).item())
P = torch.randn(700, 300, device="cuda"); Q = torch.randn(500, 300, device="cuda")
C = ops.gemm(P, Q, transb=True)
print("gemm err", ((C - P @ Q.t()).norm() / C.norm()).item())
C3 = ops.gemm_x3(P, Q, transb=True)
print("gemm_x3 err", ((C3.double() - P.double() @ Q.double().t()).norm() / C3.norm()).item())
W = ops.trtri(L[:600, :600].contiguous())
print("trtri ok", torch.isfinite(W).all().item())
torch.cuda.synchronize()
