import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
ops.TF32_MODE = "fast"
n = 8192
C = torch.randn(n, n, device="cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3
for k in (32, 128, 512):
    A = torch.randn(n, k, device="cuda")
    t1 = t(lambda: ops.gemm(A, A, transb=True, alpha=-1.0, beta=1.0, C=C))
    t0 = t(lambda: ops.gemm(A, A, transb=True, alpha=-1.0, beta=0.0, C=C))
    print(f"k={k}: beta=1 {t1:.1f} us   beta=0 {t0:.1f} us   (write-only floor 42 us, r+w floor 84 us)")
