import os, sys, torch
sys.path.insert(0, "/root/repo")
from deconditional_downscaling_b200 import ops
n, k = 8192, 128
A = torch.randn(n, k, device="cuda"); C = torch.randn(n, n, device="cuda")
Alo = ops.split_lo(A)
for beta in (1.0, 0.0):
    for _ in range(3):
        ops.gemm_x3(A, A, transb=True, alpha=-1.0, beta=beta, C=C, Alo=Alo, Blo=Alo)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm_x3(A, A, transb=True, alpha=-1.0, beta=beta, C=C, Alo=Alo, Blo=Alo); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"beta={beta}: {best*1e3:.1f} us")
ops.TF32_MODE = "fast"
for beta in (1.0, 0.0):
    for _ in range(3):
        ops.gemm(A, A, transb=True, alpha=-1.0, beta=beta, C=C)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.gemm(A, A, transb=True, alpha=-1.0, beta=beta, C=C); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"single-pass beta={beta}: {best*1e3:.1f} us")
