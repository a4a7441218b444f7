import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libdcgp.so", "libdcgp_dbg.so")
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
x = torch.randn(1024, 3, device="cuda")
A0 = ops.gram(rbf_spec(3, torch.float32, "cuda"), x, None, diag_const=0.5)
for _ in range(2):
    A = A0.clone(); ops.potrf_(A); torch.cuda.synchronize()
