import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
torch.manual_seed(0)
import itertools
for (dtype, n, m), tc in itertools.product(((torch.float32, 1024, 1024), (torch.float32, 1024, 64), (torch.float32, 1024, 256), (torch.float32, 2048, 1024), (torch.float32, 1536, 512)), (True, False)):
    ops.USE_TCGEN05 = tc
    x = torch.randn(n, 3, dtype=dtype, device="cuda")
    A = ops.gram(rbf_spec(3, dtype, "cuda"), x, None, diag_const=0.82)
    L = A.clone(); info = ops.potrf_(L); L = L.tril()
    B = torch.randn(n, m, dtype=dtype, device="cuda")
    Ld, Bd = L.double(), B.double()
    for trans in (False, True):
        X = ops.trsm_(L, B.clone(), trans=trans)
        ref = torch.linalg.solve_triangular(Ld.t() if trans else Ld, Bd, upper=trans)
        print("tc", tc, dtype, n, m, "trans", trans, "relerr", ((X.double() - ref).norm() / ref.norm()).item(), "nan", torch.isnan(X).any().item())
    X = ops.potrs_(L, B.clone())
    ref = torch.cholesky_solve(Bd, Ld)
    print("tc", tc, dtype, n, m, "potrs relerr", ((X.double() - ref).norm() / ref.norm()).item(), "info", info.item())
