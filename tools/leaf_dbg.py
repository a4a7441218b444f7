import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libdcgp.so", "libdcgp_dbg.so")
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
for dtype in (torch.float32, torch.float64):
    x = torch.randn(128, 3, dtype=dtype, device="cuda")
    A0 = ops.gram(rbf_spec(3, dtype, "cuda"), x, None, diag_const=0.5)
    for _ in range(3):
        A = A0.clone(); ops.potrf_(A); torch.cuda.synchronize()
