"""Timing + error of dcgp_trsm / dcgp_potrs: python tools/trsm_sweep.py f32|f64 n m [n m ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
dtype = torch.float32 if sys.argv[1] == "f32" else torch.float64
v = [int(a) for a in sys.argv[2:]]
def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        s, e = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        s.record(); out = fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best, out
for n, m in zip(v[::2], v[1::2]):
    torch.manual_seed(n + m)
    x = torch.randn(n, 3, dtype=dtype, device="cuda")
    A = ops.gram(rbf_spec(3, dtype, "cuda"), x, None, diag_const=0.82)
    L = A.clone(); ops.potrf_(L); L = L.tril()
    B = torch.randn(n, m, dtype=dtype, device="cuda")
    Ld, Bd = L.double(), B.double()
    out = []
    for trans in (False, True):
        t, X = timed(lambda: ops.trsm_(L, B.clone(), trans=trans))
        ref = torch.linalg.solve_triangular(Ld.t() if trans else Ld, Bd, upper=trans)
        out.append(f"trsm{'T' if trans else 'N'} {t:.3f} ms err {((X.double() - ref).norm() / ref.norm()).item():.1e}")
    t, X = timed(lambda: ops.potrs_(L, B.clone()))
    ref = torch.cholesky_solve(Bd, Ld)
    out.append(f"potrs {t:.3f} ms err {((X.double() - ref).norm() / ref.norm()).item():.1e}")
    tc, _ = timed(lambda: B.clone())
    print(f"{sys.argv[1]} n={n} m={m}: " + " | ".join(out) + f" | (clone {tc:.3f} ms)", flush=True)
    plan = ops.solve_plan(L)
    if plan is not None:
        tp, _ = timed(lambda: ops.solve_plan(L))
        t2, X = timed(lambda: ops.potrs_(L, B.clone(), plan=plan))
        print(f"    plan build {tp:.3f} ms | planned potrs {t2:.3f} ms err {((X.double() - ref).norm() / ref.norm()).item():.1e}", flush=True)
