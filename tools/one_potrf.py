"""Single dcgp_potrf (+ potrs) driver for ncu launch lists: python tools/one_potrf.py [f32|f64] [n] [m]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
dtype = torch.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else torch.float64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
x = torch.randn(n, 3, dtype=dtype, device="cuda")
A0 = ops.gram(rbf_spec(3, dtype, "cuda"), x, None, diag_const=0.82)
B0 = torch.randn(n, m, dtype=dtype, device="cuda")
for it in range(2):
    A = A0.clone(); B = B0.clone()
    torch.cuda.synchronize()
    s, e, f = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s.record(); info = ops.potrf_(A); e.record(); ops.potrs_(A, B); f.record()
    torch.cuda.synchronize()
    print(f"potrf {s.elapsed_time(e):.3f} ms  potrs {e.elapsed_time(f):.3f} ms info {info.item()}")
