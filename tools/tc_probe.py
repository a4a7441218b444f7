"""Debug probe for the tcgen05 GEMM operand majors."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
ops.TF32_MODE = "fast"
torch.manual_seed(0)
m, n, k = 1024, 2048, 64
for ta, tb in ((False, True), (False, False), (True, True), (True, False)):
    A = torch.randn((k, m) if ta else (m, k), device="cuda")
    B = torch.randn((n, k) if tb else (k, n), device="cuda")
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
    C = ops.gemm(A, B, transa=ta, transb=tb)
    err = ((C.double() - ref).norm() / ref.norm()).item()
    # diagnose: is it a permutation / partial?
    print(f"variant={os.environ.get('DCGP_TC_VARIANT','0')} ta={ta} tb={tb} relerr={err:.4f} |C|={C.norm().item():.1f} |ref|={ref.norm().item():.1f}", flush=True)
