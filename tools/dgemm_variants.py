import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import _lib
v = sys.argv[1]
if v != "0": _lib.LIB_PATH = _lib.LIB_PATH.replace("libdcgp.so", f"libdcgp_v{v}.so")
from deconditional_downscaling_b200 import ops
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda"); B = torch.randn(n, n, dtype=torch.float64, device="cuda"); C = torch.empty(n, n, dtype=torch.float64, device="cuda")
for tb, name in ((True, "NT"), (False, "NN")):
    for _ in range(2): ops.gemm(A, B, transb=tb, C=C)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3): ops.gemm(A, B, transb=tb, C=C)
    e.record(); torch.cuda.synchronize()
    print(f"variant {v} {name}: {2*n**3/(s.elapsed_time(e)/3*1e-3)/1e12:.2f} TF/s")
ref = A[:256] @ B.t()[:, :256] if False else None
