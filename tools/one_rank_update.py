"""C[n,n] -= A[n,k] A[n,k]^T (full, beta = 1) on the tcgen05 kernel, for ncu: python tools/one_rank_update.py [n] [k]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ops.TF32_MODE = "fast"
A = torch.randn(n, k, device="cuda"); C = torch.randn(n, n, device="cuda")
for _ in range(3):
    ops.gemm(A, A, transb=True, alpha=-1.0, beta=1.0, C=C)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); ops.gemm(A, A, transb=True, alpha=-1.0, beta=1.0, C=C); e.record(); torch.cuda.synchronize()
print(f"rank-{k} update of {n}^2: {s.elapsed_time(e)*1e3:.1f} us = {2*n*n*4/s.elapsed_time(e)/1e9:.2f} TB/s")
