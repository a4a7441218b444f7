import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libdcgp.so", "libdcgp_dbg.so")
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
n, m = 8192, 1024
x = torch.randn(n, 3, device="cuda")
A = ops.gram(rbf_spec(3, torch.float32, "cuda"), x, None, diag_const=0.82)
os.environ["X"] = "1"
L = A.clone()
import io, contextlib
ops.potrf_(L); L = L.tril()
B = torch.randn(n, m, device="cuda")
plan = ops.solve_plan(L)
for it in range(3):
    if it == 2: print("==== RUN", flush=True)
    ops.trsm_(L, B.clone(), plan=plan); torch.cuda.synchronize()
