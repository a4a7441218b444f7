import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libdcgp.so", "libdcgp_dbg.so")
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
dtype = torch.float32 if sys.argv[1] == "f32" else torch.float64
n = int(sys.argv[2])
x = torch.randn(n, 3, dtype=dtype, device="cuda")
A0 = ops.gram(rbf_spec(3, dtype, "cuda"), x, None, diag_const=0.82)
for it in range(3):
    A = A0.clone(); torch.cuda.synchronize()
    if it == 2: print("==== RUN", flush=True)
    ops.potrf_(A); torch.cuda.synchronize()
