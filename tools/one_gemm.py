"""Single-kernel driver for ncu captures: python tools/one_gemm.py [f64|f32] [n] [k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops  # noqa: E402

dtype = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
k = int(sys.argv[3]) if len(sys.argv) > 3 else 512
ops.TF32_MODE = "fast"
P = torch.randn(n, k, dtype=dtype, device="cuda")
C = torch.randn(n, n, dtype=dtype, device="cuda")
for _ in range(3):
    ops.gemm(P, P, transb=True, alpha=-1.0, beta=1.0, C=C, lower=True)   # SYRK trailing update shape
torch.cuda.synchronize()
print("ok")
