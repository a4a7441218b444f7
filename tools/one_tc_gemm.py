"""Single tcgen05 GEMM driver for ncu: python tools/one_tc_gemm.py [m] [n] [k] [nn|nt]   (C[m,n] = A[m,k] op(B))"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else m
k = int(sys.argv[3]) if len(sys.argv) > 3 else m
nt = (sys.argv[4] if len(sys.argv) > 4 else "nt") == "nt"
ops.TF32_MODE = "fast"
A = torch.randn(m, k, device="cuda")
B = torch.randn(n, k, device="cuda") if nt else torch.randn(k, n, device="cuda")
C = torch.empty(m, n, device="cuda")
for _ in range(3):
    ops.gemm(A, B, transb=nt, C=C)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); ops.gemm(A, B, transb=nt, C=C); e.record(); torch.cuda.synchronize()
print(f"ok {m}x{n}x{k} {'nt' if nt else 'nn'}: {s.elapsed_time(e):.4f} ms = {2*m*n*k/s.elapsed_time(e)/1e9:.1f} TF/s")
