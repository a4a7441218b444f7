import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
ops.TF32_MODE = "fast"
torch.manual_seed(0)
def t(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
for (m, n, k, ta, tb) in [(8192, 8192, 8192, False, True), (8192, 8192, 8192, False, False), (8192, 8192, 8192, True, False),
                          (7000, 9000, 1000, False, True), (8192, 8192, 128, False, True), (5000, 8192, 512, False, False)]:
    A = torch.randn((k, m) if ta else (m, k), device="cuda"); B = torch.randn((n, k) if tb else (k, n), device="cuda")
    C = ops.gemm(A, B, transa=ta, transb=tb)
    ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
    err = ((C.double() - ref).norm() / ref.norm()).item()
    ms = t(lambda: ops.gemm(A, B, transa=ta, transb=tb, C=C))
    print(f"{m}x{n}x{k} ta={int(ta)} tb={int(tb)}: err {err:.2e}  {ms*1e3:.1f} us  {2*m*n*k/ms/1e9:.1f} TF/s", flush=True)
# lower + beta + x3 path
n = 8192
P = torch.randn(n, 256, device="cuda"); Cc = torch.randn(n, n, device="cuda"); C0 = Cc.clone()
ops.gemm(P, P, transb=True, alpha=-1.0, beta=1.0, C=Cc, lower=True)
ref = (C0.double() - P.double() @ P.double().t()).tril()
print("lower beta err", ((Cc.double().tril() - ref).norm() / ref.norm()).item())
X = ops.gemm_x3(P, P, transb=True)
print("x3 err", ((X.double() - P.double() @ P.double().t()).norm() / X.norm()).item())
