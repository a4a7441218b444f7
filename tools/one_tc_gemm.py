"""Single tcgen05 GEMM driver for ncu: python tools/one_tc_gemm.py [n] [k] [x3]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
ops.TF32_MODE = "fast"
A = torch.randn(n, k, device="cuda"); B = torch.randn(n, k, device="cuda"); C = torch.empty(n, n, device="cuda")
for _ in range(3):
    ops.gemm(A, B, transb=True, C=C)
torch.cuda.synchronize()
print("ok")
