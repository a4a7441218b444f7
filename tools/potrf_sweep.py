"""Timing + residual of dcgp_potrf over sizes: python tools/potrf_sweep.py f32|f64 n [n ...]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
dtype = torch.float32 if sys.argv[1] == "f32" else torch.float64
for n in [int(a) for a in sys.argv[2:]]:
    x = torch.randn(n, 3, dtype=dtype, device="cuda")
    A0 = ops.gram(rbf_spec(3, dtype, "cuda"), x, None, diag_const=0.82)
    best = 1e9
    for it in range(4):
        A = A0.clone()
        torch.cuda.synchronize()
        s, e = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        t0 = time.perf_counter(); s.record(); info = ops.potrf_(A); e.record(); host = (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    L = A.tril().double()
    v = torch.randn(n, 4, dtype=torch.float64, device="cuda")
    ref = A0.double() @ v
    err = ((L @ (L.t() @ v) - ref).norm() / ref.norm()).item()
    print(f"{sys.argv[1]} n={n} potrf {best:.3f} ms  {n**3/3/best/1e9:.2f} TF/s  host-issue {host:.2f} ms info {info.item()} recon_err {err:.2e}", flush=True)
