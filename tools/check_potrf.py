"""Factor / solve / invert a few SPD matrices with the library as configured by the environment and compare with LAPACK:
python tools/check_potrf.py [f32|f64|both] [sizes...]   (used for A/B runs of the DCGP_* switches)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "both"
sizes = [int(a) for a in sys.argv[2:]] or [7, 16, 100, 128, 129, 200, 256, 300, 517, 1024, 2048, 4096]
torch.manual_seed(0)
ok = True
for dt, tol in ((torch.float64, 1e-10), (torch.float32, 2e-5)):
    if which not in ("both", "f32" if dt == torch.float32 else "f64"):
        continue
    for n in sizes:
        M = torch.randn(n, n, dtype=torch.float64, device="cuda")
        A = (M @ M.t() / n + torch.eye(n, dtype=torch.float64, device="cuda")).to(dt)
        L = A.clone()
        info = ops.potrf_(L)
        L = L.tril()
        ref = torch.linalg.cholesky(A.double())
        err = ((L.double() - ref).norm() / ref.norm()).item()
        B = torch.randn(n, 37, dtype=dt, device="cuda")
        X = B.clone()
        ops.potrs_(L, X)
        res = ((A.double() @ X.double() - B.double()).norm() / B.double().norm()).item()
        good = int(info.item()) == 0 and err < tol and res < 50 * tol
        ok &= good
        print(f"{str(dt)[6:]} n={n:5d} info {int(info.item())} factor err {err:.2e} solve resid {res:.2e} {'ok' if good else 'FAIL'}", flush=True)
    # a matrix that is not positive definite: the first bad leading minor must be reported
    n = 600
    A = torch.eye(n, dtype=dt, device="cuda") * 2
    A[411, 411] = -1.0
    info = ops.potrf_(A.clone())
    good = int(info.item()) == 412
    ok &= good
    print(f"{str(dt)[6:]} not-PD info {int(info.item())} (want 412) {'ok' if good else 'FAIL'}")
print("ALL OK" if ok else "FAILED")
