"""C[n,n] -= A[n,k] A[n,k]^T (full, beta = 1) as the factorisation issues it: 3xTF32 on the tcgen05 kernel with remainder
operands (fat stages).  python tools/one_rank_update_x3.py [n] [k] [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
A = torch.randn(n, k, device="cuda"); C = torch.randn(n, n, device="cuda")
Alo = ops.split_lo(A)
for _ in range(3):
    ops.gemm_x3(A, A, transb=True, alpha=-1.0, beta=1.0, C=C, Alo=Alo, Blo=Alo)
torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.gemm_x3(A, A, transb=True, alpha=-1.0, beta=1.0, C=C, Alo=Alo, Blo=Alo); e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
print(f"3xTF32 rank-{k} update of {n}^2: {best*1e3:.1f} us = {2*n*n*4/best/1e9:.2f} TB/s of C traffic")
